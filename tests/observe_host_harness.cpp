// CPU harness for tests/test_observe_host.py: the statements each warp of k_observe runs for its env
// (hope_b200/csrc/observe_body.inc: ego-frame edge staging, 120-ray cast, action-mask sweep, 5-tap post-process) compiled
// with g++ on the 32-fiber warp emulation of tests/warp_emu.h, so the lidar and action-mask traces recorded from the
// unmodified reference (tests/golden/episodes_*.npz) can be replayed through the product's own code without a GPU.
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
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
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

namespace hope {
#include "../hope_b200/csrc/hope_types.cuh"
#include "../hope_b200/csrc/observe.cuh"

static void observe_one(const int env, const int lane, ObserveSmem &sm, Pool pool, EnvState st, Tables tb, hope_params par, hope_out out) {
#include "../hope_b200/csrc/observe_body.inc"
}
}  // namespace hope

// One env.  pose = (x, y, heading); obs[MAXO][4][2], nv[MAXO] with the rings compacted to the front (nobs of them).
// tables: ray_a[120] ray_b[120] lidar_base[120] mask_base[120] w_lo[10] w_hi[10] pmaxk[1200][10][42] pmax[1200] gpmax[120]
// (the last three are what k_table_reduce / k_table_group derive from dist_star at upload).
// Outputs: lidar[120], mask[42], steps[42].  Returns 0, or -2 for a convergence error of the warp.
extern "C" int observe_host(const double *pose, int nobs, const double *obs, const uint8_t *nv, const double *ray_a, const double *ray_b,
                            const double *lidar_base, const double *mask_base, const double *w_lo, const double *w_hi, const double *pmaxk,
                            const double *pmax, const double *gpmax, double lidar_range, double *lidar, double *mask, uint8_t *steps) {
    using namespace hope;
    static ObserveSmem sm;
    alignas(16) static double verts[MAXE * 2];
    std::memcpy(verts, obs, sizeof(verts));
    double cs[2];
    sincos(pose[2], &cs[1], &cs[0]);  // k_advance hands cos / sin of the heading to k_observe (EnvState::cs)
    int scene = 0;
    Pool pool{};
    pool.obs = verts; pool.nv = nv; pool.nobs = &nobs; pool.size = 1;
    EnvState st{};
    st.pose = const_cast<double *>(pose); st.cs = cs; st.scene = &scene;
    Tables tb{};
    tb.ray_a = ray_a; tb.ray_b = ray_b; tb.lidar_base = lidar_base; tb.mask_base = mask_base; tb.w_lo = w_lo; tb.w_hi = w_hi;
    tb.pmaxk = pmaxk; tb.pmax = pmax; tb.gpmax = gpmax;
    hope_params par;
    std::memset(&par, 0, sizeof(par));
    par.lidar_range = lidar_range;
    hope_out out;
    std::memset(&out, 0, sizeof(out));
    out.lidar = lidar; out.mask = mask; out.mask_steps = steps;
    const char *err = warp_emu::run([&](int lane) { observe_one(0, lane, sm, pool, st, tb, par, out); });
    if (err) {
        if (getenv("WARP_EMU_VERBOSE")) fprintf(stderr, "warp_emu: %s\n", err);
        return -2;
    }
    return 0;
}
