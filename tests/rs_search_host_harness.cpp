// CPU harness for tests/test_rs_search_host.py: the whole Reeds-Shepp search of one env the way the kernels run it —
// k_rs_enumerate's enumerate_env (rs_enumerate.cuh: admitted words in heapdict pop order with the 1.6 x cut-off),
// k_rs_walk's plan_word (rs_walk.cuh), k_rs_check's warp code (rs_check.cuh, on the 32-fiber warp emulation of
// tests/warp_emu.h) for EVERY tried word, and k_rs_select's body (rs_select_body.inc: first clean word wins) — so the
// find_rs_path results recorded from the unmodified reference (car_parking_base.py:413-450; tests/golden/episodes*_*.npz:
// found, words tried, candidates, segment types and lengths) can be replayed through the product's own code without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)
#define __restrict__
#define HOPE_CONSTANT static const
#define HOPE_STAT(i, v) ((void)0)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

#include "warp_emu.h"

#include "../include/hope_b200.h"
#include "../hope_b200/csrc/hope_device.cuh"

#ifndef HOPE_CHK_EDGE_EXIT
#define HOPE_CHK_EDGE_EXIT 1
#endif

namespace hope {
#include "../hope_b200/csrc/hope_types.cuh"
static inline double4 ld_aabb(const double4 *p) { return *p; }  // the kernel's version is two 128-bit read-only loads
#include "../hope_b200/csrc/div_pair.cuh"
#include "../hope_b200/csrc/rs_words.cuh"
#include "../hope_b200/csrc/rs_enumerate.cuh"
#include "../hope_b200/csrc/rs_walk.cuh"
#include "../hope_b200/csrc/rs_check.cuh"

static void select_one(const int i, Tables tb, RsScratch rs, hope_out out) {
#include "../hope_b200/csrc/rs_select_body.inc"
}
}  // namespace hope

// One env.  pose / dest = (x, y, heading); gate = k_advance's RS gate of this step (t > 1, CONTINUE, closer than 10 m);
// obs[MAXO][4][2], nv[MAXO] compacted to the front.  Outputs as in hope_out.  Returns 0, -2 on a warp convergence error.
extern "C" int rs_search_host(const double *pose, const double *dest, const double *bounds, int gate, int nobs, const double *obs,
                              const uint8_t *nv, const double *box_x, const double *box_y, double maxc, double rs_step, uint8_t *found,
                              uint8_t *nseg, uint8_t *types /*5*/, double *lengths /*5*/, double *L, uint8_t *ncand, uint8_t *ntried) {
    using namespace hope;
    static double meta[META];
    static double4 aabb[MAXO];
    alignas(16) static double verts[MAXE * 2];
    static uint8_t nvs[MAXO];
    std::memset(meta, 0, sizeof(meta)); std::memset(aabb, 0, sizeof(aabb)); std::memset(nvs, 0, sizeof(nvs));
    std::memcpy(verts, obs, sizeof(verts));
    for (int k = 0; k < 3; ++k) meta[M_DEST + k] = dest[k];
    for (int k = 0; k < 4; ++k) meta[M_BOUNDS + k] = bounds[k];
    for (int o = 0; o < nobs; ++o) {
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        nvs[o] = nv[o];
        for (int v = 0; v < nv[o]; ++v) {
            const double x = obs[(o * MAXV + v) * 2], y = obs[(o * MAXV + v) * 2 + 1];
            xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        }
        aabb[o] = make_double4(xmin, xmax, ymin, ymax);
    }
    double cs[2];
    sincos(pose[2], &cs[1], &cs[0]);
    int scene = 0;
    uint8_t gate8 = gate ? 1 : 0;
    unsigned long long counters[8] = {0};
    Pool pool{verts, nvs, reinterpret_cast<const double *>(aabb), meta, &nobs, 1};
    EnvState st{};
    st.pose = const_cast<double *>(pose); st.cs = cs; st.scene = &scene; st.gate = &gate8; st.counters = counters;
    Tables tb{};
    tb.maxc = maxc;
    static RsWord words[MAXW];
    static WordSlot slot;
    uint8_t r_ntry = 0, r_ncand = 0, item_bad[MAXW];
    int item_base = 0;
    RsScratch rs{};
    rs.words = words; rs.ntry = &r_ntry; rs.ncand = &r_ncand; rs.item_base = &item_base; rs.item_bad = item_bad;
    hope_out out;
    std::memset(&out, 0, sizeof(out));
    out.rs_found = found; out.rs_nseg = nseg; out.rs_types = types; out.rs_lengths = lengths; out.rs_L = L; out.rs_ncand = ncand; out.rs_ntried = ntried;
    hope_params par;
    std::memset(&par, 0, sizeof(par));
    for (int k = 0; k < 4; ++k) { par.box_x[k] = box_x[k]; par.box_y[k] = box_y[k]; }
    par.rs_step = rs_step;

    const int ntry = enumerate_env(0, pool, st, tb, rs, out);                 // k_rs_enumerate
    CheckEnv E;
    E.q0x = pose[0]; E.q0y = pose[1]; E.q0h = pose[2];
    E.cg = cs[0]; E.sg = -cs[1];
    E.xmin = bounds[0]; E.xmax = bounds[1]; E.ymin = bounds[2]; E.ymax = bounds[3];
    E.maxc = maxc; E.step = rs_step * maxc;
    E.nobs = nobs; E.aabb = aabb; E.verts = reinterpret_cast<const double2 *>(verts); E.nvp = nvs;
    for (int k = 0; k < ntry; ++k) {
        plan_word(slot, words[k], maxc, rs_step * maxc);                       // k_rs_walk
        int verdict[32];
        const char *err = warp_emu::run([&](int lane) {                        // k_rs_check
            bool bad = false;
            int chunk_base = 0;
            for (;;) {
                bad = chunk_is_bad(slot, E, par, lane);
                if (bad || slot.total >= 0) break;
                chunk_base += RS_CHUNK;
                __syncwarp();
                if (lane == 0) walk_chunk(slot, slot.len, E.step, chunk_base);
                __syncwarp();
            }
            verdict[lane] = bad ? 1 : 0;
        });
        if (err) {
            if (getenv("WARP_EMU_VERBOSE")) fprintf(stderr, "warp_emu: %s\n", err);
            return -2;
        }
        item_bad[k] = (uint8_t)verdict[0];
    }
    select_one(0, tb, rs, out);                                                // k_rs_select
    return 0;
}
