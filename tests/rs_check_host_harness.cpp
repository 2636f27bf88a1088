// CPU harness for tests/test_rs_check_host.py: the warp-level trajectory check of k_rs_check
// (hope_b200/csrc/rs_check.cuh: chunk_is_bad / warp_samples_hit, with rs_walk.cuh's plan and walker and div_pair.cuh)
// compiled with g++ and run on the 32-fiber warp emulation of tests/warp_emu.h, so that is_traj_valid verdicts
// recorded from the unmodified reference (tests/golden/traj_valid.npz) can be replayed through the product's own code
// without a GPU — including the trailing-zero path (reeds_shepp.py:501-505) that no generated
// scene reaches.  Built twice by the test: -DHOPE_CHK_EDGE_EXIT=1 (the shipped vote placement) and =0.
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
#define HOPE_CONSTANT static const
#define HOPE_STAT(i, v) ((void)0)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
struct double2 { double x, y; };
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
#ifndef HOPE_CHK_POOLED
#define HOPE_CHK_POOLED 1
#endif

namespace hope {
constexpr int MAXW = 16;  // as in hope_kernels.cu
constexpr int MAXO = HOPE_MAX_OBS;
constexpr int MAXV = HOPE_MAX_VERTS;
static inline double4 ld_aabb(const double4 *p) { return *p; }  // the kernel's version is two 128-bit read-only loads
#include "../hope_b200/csrc/div_pair.cuh"
#include "../hope_b200/csrc/rs_words.cuh"
#include "../hope_b200/csrc/rs_walk.cuh"
#include "../hope_b200/csrc/rs_check.cuh"
#include "../hope_b200/csrc/rs_check_pooled.cuh"
}  // namespace hope

// One tried word ready for the check: its sampling plan (k_rs_walk's output) and the check environment k_rs_check builds.
struct Prepared {
    hope::WordSlot s;
    hope::CheckEnv E;
    double4 aabb[hope::MAXO];
    double2 verts[hope::MAXO * hope::MAXV];
    uint8_t nvs[hope::MAXO];
};

// q = (start x, y, heading, goal x, y, heading); word = index in the product's enumeration order (== the reference's
// calc_all_paths order).  obs[nobs][4][2], nv[nobs].
static int prepare(Prepared &P, const double *q, double maxc, double rs_step, int word, const double *bounds, int nobs, const double *obs,
                   const uint8_t *nv, int force_zero_tail) {
    using namespace hope;
    WordList w;
    enumerate_words(q[0], q[1], q[2], q[3], q[4], q[5], maxc, w, nullptr);
    if (word < 0 || word >= w.count || nobs > MAXO) return -1;
    RsWord rw;
    std::memset(&rw, 0, sizeof(rw));
    for (int i = 0; i < 5; ++i) { rw.len[i] = w.len[word][i]; rw.types[i] = (uint8_t)((w.ty[word] >> (4 * i)) & 0xF); }
    rw.L = w.L[word]; rw.n = w.n[word];
    plan_word(P.s, rw, maxc, rs_step * maxc);
    if (force_zero_tail) P.s.end_lx = 0.0;
    std::memset(P.aabb, 0, sizeof(P.aabb)); std::memset(P.verts, 0, sizeof(P.verts)); std::memset(P.nvs, 0, sizeof(P.nvs));
    for (int o = 0; o < nobs; ++o) {  // per-ring bounding boxes as hope_set_scene_pool stores them (xmin xmax ymin ymax)
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        P.nvs[o] = nv[o];
        for (int v = 0; v < MAXV; ++v) {
            const double x = obs[(o * MAXV + v) * 2], y = obs[(o * MAXV + v) * 2 + 1];
            P.verts[o * MAXV + v] = make_double2(x, y);
            if (v < nv[o]) { xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y); }
        }
        P.aabb[o] = make_double4(xmin, xmax, ymin, ymax);
    }
    double sh, ch;
    sincos(q[2], &sh, &ch);  // k_advance hands cos / sin of the heading to the check (EnvState::cs)
    CheckEnv &E = P.E;
    E.q0x = q[0]; E.q0y = q[1]; E.q0h = q[2];
    E.cg = ch; E.sg = -sh;
    E.xmin = bounds[0]; E.xmax = bounds[1]; E.ymin = bounds[2]; E.ymax = bounds[3];
    E.maxc = maxc; E.step = rs_step * maxc;
    E.nobs = nobs; E.aabb = P.aabb; E.verts = P.verts; E.nvp = P.nvs;
    return 0;
}

static hope_params make_params(const double *box_x, const double *box_y, double rs_step) {
    hope_params par;
    std::memset(&par, 0, sizeof(par));
    for (int k = 0; k < 4; ++k) { par.box_x[k] = box_x[k]; par.box_y[k] = box_y[k]; }
    par.rs_step = rs_step;
    return par;
}

static const char *report(const char *err) {
    if (err && getenv("WARP_EMU_VERBOSE")) fprintf(stderr, "warp_emu: %s\n", err);
    return err;
}

// the word loop body of k_rs_check on one word, one fiber per lane
static int whole_warp_verdict(Prepared &P, const hope_params &par, unsigned long long *n_collectives) {
    using namespace hope;
    int verdict[32];
    const char *err = warp_emu::run([&](int lane) {
        bool bad = false;
        int chunk_base = 0;
        for (;;) {
#if HOPE_CHK_POOLED
            static CheckSmem cs;
            bad = chunk_is_bad_pooled(P.s, P.E, par, lane, cs);
#else
            bad = chunk_is_bad(P.s, P.E, par, lane);
#endif
            if (bad || P.s.total >= 0) break;
            chunk_base += RS_CHUNK;
            __syncwarp();
            if (lane == 0) walk_chunk(P.s, P.s.len, P.E.step, chunk_base);
            __syncwarp();
        }
        verdict[lane] = bad ? 1 : 0;
    }, n_collectives);
    if (report(err)) return -2;
    for (int l = 1; l < 32; ++l)
        if (verdict[l] != verdict[0]) return -3;
    return verdict[0];
}

// Returns 1 = the word leaves the map or touches an obstacle, 0 = clean, < 0 = error (-1 bad word index, -2 convergence
// error of the warp, -3 lanes disagree).
extern "C" int rs_check_host(const double *q, double maxc, double rs_step, int word, const double *bounds, int nobs, const double *obs,
                             const uint8_t *nv, const double *box_x, const double *box_y, int force_zero_tail, int *n_samples,
                             unsigned long long *n_collectives) {
    static Prepared P;
    if (prepare(P, q, maxc, rs_step, word, bounds, nobs, obs, nv, force_zero_tail)) return -1;
    const hope_params par = make_params(box_x, box_y, rs_step);
    const int v = whole_warp_verdict(P, par, n_collectives);
    if (n_samples) *n_samples = P.s.total;
    return v;
}
