// CPU harness for tests/test_advance_host.py: the statements each thread of k_advance runs (hope_b200/csrc/advance_body.inc
// with advance.cuh's warp-pooled collision test) compiled with g++ on the 32-fiber warp emulation of tests/warp_emu.h.
// One emulated warp = 32 envs in lock step, exactly like on the GPU, so the pose / status / reward / substep traces
// recorded from the unmodified reference (tests/golden/episodes_*.npz) can be replayed through the product's own code
// without a GPU.  The scene packing below restates what hope_set_scene_pool does on the host (that code needs a CUDA
// context); it is test plumbing, the thing under test is the kernel body.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __double2float_rd(double a) { float f = (float)a; return (double)f > a ? std::nextafterf(f, -INFINITY) : f; }
static inline float __double2float_ru(double a) { float f = (float)a; return (double)f < a ? std::nextafterf(f, INFINITY) : f; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

#include "warp_emu.h"

#include "../include/hope_b200.h"
#include "../hope_b200/csrc/hope_device.cuh"

namespace hope {
#include "../hope_b200/csrc/hope_types.cuh"
static inline double4 ld_aabb(const double4 *p) { return *p; }  // the kernel's version is two 128-bit read-only loads
#include "../hope_b200/csrc/advance.cuh"

static void advance_one(const int n, const int gi, const int lane, AdvanceSmem &sm, Pool pool, EnvState st, const double *action, hope_params par,
                        hope_out out, int reset_all, int reset_stride, bool raw_action = false) {
#include "../hope_b200/csrc/advance_body.inc"
}
}  // namespace hope

namespace {
struct Sim {
    int n = 0;
    hope_params par;
    std::vector<double> obs, aabb, meta, pose, cs, accum, traj;
    std::vector<uint8_t> nv, pending, gate;
    std::vector<int> nobs, t, scene, traj_n;
    unsigned long long counters[8] = {0};
    // outputs
    std::vector<double> o_pose, o_target, o_reward, o_reward_info;
    std::vector<int32_t> o_status;
    std::vector<uint8_t> o_done, o_substeps, o_retreated, o_was_reset;
} g;
}  // namespace

extern "C" int adv_create(int n, const hope_params *par) {
    using namespace hope;
    g = Sim();
    g.n = n; g.par = *par;
    g.par.auto_reset = 0; g.par.regen_on_reset = 0;
    g.obs.assign((size_t)n * MAXE * 2, 0.0); g.aabb.assign((size_t)n * MAXO * 4, 0.0); g.meta.assign((size_t)n * META, 0.0);
    g.nv.assign((size_t)n * MAXO, 0); g.nobs.assign(n, 0);
    g.pose.assign(3 * n, 0.0); g.cs.assign(2 * n, 0.0); g.accum.assign(n, 0.0); g.t.assign(n, 0); g.scene.resize(n);
    for (int i = 0; i < n; ++i) g.scene[i] = i;
    g.pending.assign(n, 0); g.gate.assign(n, 0); g.traj.assign((size_t)n * 80, 0.0); g.traj_n.assign(n, 0);
    g.o_pose.assign(3 * n, 0.0); g.o_target.assign(5 * n, 0.0); g.o_reward.assign(n, 0.0); g.o_reward_info.assign(5 * n, 0.0);
    g.o_status.assign(n, 0); g.o_done.assign(n, 0); g.o_substeps.assign(n, 0); g.o_retreated.assign(n, 0); g.o_was_reset.assign(n, 0);
    return 0;
}

// scene i of the pool (= env i): start[3], dest[3], bounds[4], obs[MAXO][4][2], nverts[MAXO]
extern "C" int adv_set_scene(int i, const double *start, const double *dest, const double *bounds, const double *obs_xy, const int32_t *nverts) {
    using namespace hope;
    const hope_params &par = g.par;
    double *m = &g.meta[(size_t)i * META];
    for (int k = 0; k < 3; ++k) { m[M_START + k] = start[k]; m[M_DEST + k] = dest[k]; }
    for (int k = 0; k < 4; ++k) m[M_BOUNDS + k] = bounds[k];
    const double c = cos(dest[2]), s = sin(dest[2]), ms = -s;
    double bx[4], by[4];
    for (int k = 0; k < 4; ++k) {
        bx[k] = c * par.box_x[k] + ms * par.box_y[k] + dest[0];
        by[k] = s * par.box_x[k] + c * par.box_y[k] + dest[1];
        m[M_DBX + k] = bx[k]; m[M_DBY + k] = by[k];
    }
    double sa = 0.0;
    for (int k = 0; k < 4; ++k) { const int j = (k + 1) & 3; sa += bx[k] * by[j] - bx[j] * by[k]; }
    m[M_DAREA] = fabs(sa) * 0.5;
    m[M_DNORM] = fmax(hypot(dest[0] - start[0], dest[1] - start[1]), 10.0);
    m[M_DAABB] = fmin(fmin(bx[0], bx[1]), fmin(bx[2], bx[3])); m[M_DAABB + 1] = fmax(fmax(bx[0], bx[1]), fmax(bx[2], bx[3]));
    m[M_DAABB + 2] = fmin(fmin(by[0], by[1]), fmin(by[2], by[3])); m[M_DAABB + 3] = fmax(fmax(by[0], by[1]), fmax(by[2], by[3]));
    int no = 0;
    for (int k = 0; k < MAXO; ++k) {
        const int v = nverts[k];
        if (v == 0) continue;
        if (v < 3 || v > MAXV) return -1;
        const double *src = obs_xy + (size_t)k * MAXV * 2;
        double *dst = &g.obs[((size_t)i * MAXO + no) * MAXV * 2];
        double xmn = src[0], xmx = src[0], ymn = src[1], ymx = src[1];
        for (int j = 0; j < v; ++j) {
            dst[2 * j] = src[2 * j]; dst[2 * j + 1] = src[2 * j + 1];
            xmn = fmin(xmn, src[2 * j]); xmx = fmax(xmx, src[2 * j]); ymn = fmin(ymn, src[2 * j + 1]); ymx = fmax(ymx, src[2 * j + 1]);
        }
        double *bb = &g.aabb[((size_t)i * MAXO + no) * 4];
        bb[0] = xmn; bb[1] = xmx; bb[2] = ymn; bb[3] = ymx;
        g.nv[(size_t)i * MAXO + no] = (uint8_t)v;
        ++no;
    }
    g.nobs[i] = no;
    return 0;
}

// One launch of k_advance over all envs: action == NULL or reset_all -> the reset step (car_parking_base.py:127-138).
extern "C" int adv_launch(const double *action, int reset_all) {
    using namespace hope;
    Pool pool{g.obs.data(), g.nv.data(), g.aabb.data(), g.meta.data(), g.nobs.data(), g.n};
    EnvState st{g.pose.data(), g.cs.data(), g.t.data(), g.accum.data(), g.scene.data(), g.pending.data(), g.gate.data(), g.counters,
                g.traj.data(), g.traj_n.data()};
    hope_out out;
    std::memset(&out, 0, sizeof(out));
    out.pose = g.o_pose.data(); out.target = g.o_target.data(); out.reward = g.o_reward.data(); out.reward_info = g.o_reward_info.data();
    out.status = g.o_status.data(); out.done = g.o_done.data(); out.substeps = g.o_substeps.data(); out.retreated = g.o_retreated.data();
    out.was_reset = g.o_was_reset.data();
    static AdvanceSmem sm;
    for (int base = 0; base < g.n; base += 32) {
        const char *err = warp_emu::run([&](int lane) { advance_one(g.n, base + lane, lane, sm, pool, st, action, g.par, out, reset_all, g.n); });
        if (err) {
            if (getenv("WARP_EMU_VERBOSE")) fprintf(stderr, "warp_emu: %s\n", err);
            return -2;
        }
    }
    return 0;
}

extern "C" void adv_read(double *pose, int32_t *status, double *reward, double *reward_info, double *target, uint8_t *substeps, uint8_t *retreated,
                         uint8_t *done, unsigned long long *exact_fallbacks) {
    const int n = g.n;
    std::memcpy(pose, g.o_pose.data(), sizeof(double) * 3 * n); std::memcpy(status, g.o_status.data(), sizeof(int32_t) * n);
    std::memcpy(reward, g.o_reward.data(), sizeof(double) * n); std::memcpy(reward_info, g.o_reward_info.data(), sizeof(double) * 5 * n);
    std::memcpy(target, g.o_target.data(), sizeof(double) * 5 * n); std::memcpy(substeps, g.o_substeps.data(), n);
    std::memcpy(retreated, g.o_retreated.data(), n); std::memcpy(done, g.o_done.data(), n);
    if (exact_fallbacks) *exact_fallbacks = g.counters[2];
}
