// CPU harness for tests/test_step_host.py: the whole env step as the kernels run it, chained on the host for test purposes —
// k_advance (advance_body.inc, 32 envs per emulated warp), k_observe (observe_body.inc, one emulated warp per env),
// k_rs_enumerate's enumerate_env, k_rs_walk's plan_word, k_rs_check (rs_check_pooled.cuh, or with -DHOPE_CHK_POOLED=0 the per-lane edge
// loop of rs_check.cuh) and k_rs_select (rs_select_body.inc) — with
// the state arrays the kernels keep between steps.  tests/test_step_host.py runs it in lock step with the C oracle on
// generated scenes, the way tests/test_gpu_parity.py does with the real kernels on the GPU.  This is test plumbing around
// the product's device source, not a CPU path of the product (nothing under hope_b200/ can reach it).
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
static inline double __drcp_rn(double a);  // below: counts one unit of "division" work per call
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
static inline double __drcp_rn(double a) { warp_emu::work(0, 1); return 1.0 / a; }
#define sincos(a, s, c) (warp_emu::work(1, 1), ::sincos(a, s, c))


#include "../include/hope_b200.h"
#include "../hope_b200/csrc/hope_device.cuh"

#ifndef HOPE_CHK_EDGE_EXIT
#define HOPE_CHK_EDGE_EXIT 1
#endif
#ifndef HOPE_CHK_POOLED
#define HOPE_CHK_POOLED 1
#endif

namespace hope {
#include "../hope_b200/csrc/hope_types.cuh"
static inline double4 ld_aabb(const double4 *p) { return *p; }  // the kernel's version is two 128-bit read-only loads
#include "../hope_b200/csrc/advance.cuh"
#include "../hope_b200/csrc/observe.cuh"
#include "../hope_b200/csrc/rs_words.cuh"
#include "../hope_b200/csrc/rs_enumerate.cuh"
#include "../hope_b200/csrc/rs_walk.cuh"
#include "../hope_b200/csrc/rs_check.cuh"
#include "../hope_b200/csrc/rs_check_pooled.cuh"

static void advance_one(const int n, const int gi, const int lane, AdvanceSmem &sm, Pool pool, EnvState st, const double *action, hope_params par,
                        hope_out out, int reset_all, int reset_stride, bool raw_action = false) {
#include "../hope_b200/csrc/advance_body.inc"
}
static void observe_one(const int env, const int lane, ObserveSmem &sm, Pool pool, EnvState st, Tables tb, hope_params par, hope_out out) {
#include "../hope_b200/csrc/observe_body.inc"
}
static void select_one(const int i, Tables tb, RsScratch rs, hope_out out) {
#include "../hope_b200/csrc/rs_select_body.inc"
}
}  // namespace hope

namespace {
using namespace hope;
struct Sim {
    int n = 0;
    hope_params par;
    double maxc = 0.0;
    std::vector<double> obs, aabb, meta, pose, cs, accum, traj, tab;
    std::vector<uint8_t> nv, pending, gate;
    std::vector<int> nobs, t, scene, traj_n;
    unsigned long long counters[8] = {0};
    // RS scratch
    std::vector<RsWord> words;
    std::vector<uint8_t> ntry, ncand, item_bad;
    std::vector<int> item_base, items;
    // outputs
    std::vector<double> o_pose, o_target, o_reward, o_reward_info, o_lidar, o_mask, o_rs_lengths, o_rs_L;
    std::vector<int32_t> o_status;
    std::vector<uint8_t> o_done, o_substeps, o_retreated, o_was_reset, o_mask_steps, o_rs_found, o_rs_nseg, o_rs_types, o_rs_ncand, o_rs_ntried;
} g;

CheckEnv check_env(int env) {  // what k_rs_check loads for a work item of env `env`
    const int sid = g.scene[env];
    const double *meta = &g.meta[(size_t)sid * META];
    CheckEnv E;
    E.q0x = g.pose[3 * env]; E.q0y = g.pose[3 * env + 1]; E.q0h = g.pose[3 * env + 2];
    E.cg = g.cs[2 * env]; E.sg = -g.cs[2 * env + 1];
    E.xmin = meta[M_BOUNDS]; E.xmax = meta[M_BOUNDS + 1]; E.ymin = meta[M_BOUNDS + 2]; E.ymax = meta[M_BOUNDS + 3];
    E.maxc = g.maxc; E.step = g.par.rs_step * g.maxc;
    E.nobs = g.nobs[sid];
    E.aabb = reinterpret_cast<const double4 *>(g.aabb.data()) + (size_t)sid * MAXO;
    E.verts = reinterpret_cast<const double2 *>(g.obs.data()) + (size_t)sid * MAXE;
    E.nvp = g.nv.data() + (size_t)sid * MAXO;
    return E;
}
}  // namespace

// tables: ray_a[120] ray_b[120] lidar_base[120] mask_base[120] w_lo[10] w_hi[10] pmaxk[1200*10*42] pmax[1200] gpmax[120]
extern "C" int step_create(int n, const hope_params *par, const double *ray_a, const double *ray_b, const double *lidar_base, const double *mask_base,
                           const double *w_lo, const double *w_hi, const double *pmaxk, const double *pmax, const double *gpmax) {
    g = Sim();
    g.n = n; g.par = *par;
    g.par.auto_reset = 0; g.par.regen_on_reset = 0;
    g.maxc = tan(g.par.valid_steer[1]) / g.par.wheel_base;
    g.obs.assign((size_t)n * MAXE * 2, 0.0); g.aabb.assign((size_t)n * MAXO * 4 + 2, 0.0); g.meta.assign((size_t)n * META, 0.0);
    g.nv.assign((size_t)n * MAXO, 0); g.nobs.assign(n, 0);
    g.pose.assign(3 * n, 0.0); g.cs.assign(2 * n, 0.0); g.accum.assign(n, 0.0); g.t.assign(n, 0); g.scene.resize(n);
    for (int i = 0; i < n; ++i) g.scene[i] = i;
    g.pending.assign(n, 0); g.gate.assign(n, 0); g.traj.assign((size_t)n * 80, 0.0); g.traj_n.assign(n, 0);
    const size_t NK = (size_t)NUP * NITER * NACT;
    g.tab.assign(512 + NK + NUP + NRAY, 0.0);  // ray_a ray_b lidar_base mask_base w_lo w_hi | pmaxk | pmax | gpmax
    std::memcpy(&g.tab[0], ray_a, 8 * NRAY); std::memcpy(&g.tab[120], ray_b, 8 * NRAY); std::memcpy(&g.tab[240], lidar_base, 8 * NRAY);
    std::memcpy(&g.tab[360], mask_base, 8 * NRAY); std::memcpy(&g.tab[480], w_lo, 8 * 10); std::memcpy(&g.tab[496], w_hi, 8 * 10);
    std::memcpy(&g.tab[512], pmaxk, 8 * NK); std::memcpy(&g.tab[512 + NK], pmax, 8 * NUP); std::memcpy(&g.tab[512 + NK + NUP], gpmax, 8 * NRAY);
    g.words.resize((size_t)n * MAXW); g.ntry.assign(n, 0); g.ncand.assign(n, 0); g.item_bad.assign((size_t)n * MAXW, 0);
    g.item_base.assign(n, 0); g.items.assign((size_t)n * MAXW, 0);
    g.o_pose.assign(3 * n, 0.0); g.o_target.assign(5 * n, 0.0); g.o_reward.assign(n, 0.0); g.o_reward_info.assign(5 * n, 0.0);
    g.o_lidar.assign((size_t)NRAY * n, 0.0); g.o_mask.assign((size_t)NACT * n, 0.0); g.o_rs_lengths.assign(5 * n, 0.0); g.o_rs_L.assign(n, 0.0);
    g.o_status.assign(n, 0); g.o_done.assign(n, 0); g.o_substeps.assign(n, 0); g.o_retreated.assign(n, 0); g.o_was_reset.assign(n, 0);
    g.o_mask_steps.assign((size_t)NACT * n, 0); g.o_rs_found.assign(n, 0); g.o_rs_nseg.assign(n, 0); g.o_rs_types.assign(5 * n, 0);
    g.o_rs_ncand.assign(n, 0); g.o_rs_ntried.assign(n, 0);
    return 0;
}

// scene i of the pool (= env i), packed the way hope_set_scene_pool packs it on the host
extern "C" int step_set_scene(int i, const double *start, const double *dest, const double *bounds, const double *obs_xy, const int32_t *nverts) {
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

static unsigned long long g_check_votes = 0, g_check_runs = 0;  // warp-level votes / warp runs of the k_rs_check stage
static unsigned long long g_check_div = 0, g_check_sincos = 0;  // warp-level cost model of the check stage (warp_emu::work)
extern "C" void step_check_work(unsigned long long *votes, unsigned long long *runs) { *votes = g_check_votes; *runs = g_check_runs; }
extern "C" void step_check_cost(unsigned long long *div, unsigned long long *sc) { *div = g_check_div; *sc = g_check_sincos; }

static int fail(const char *err) {
    if (getenv("WARP_EMU_VERBOSE")) fprintf(stderr, "warp_emu: %s\n", err);
    return -2;
}

// One step of all envs (hope_step with HOPE_STAGE_ALL): action == NULL / reset_all -> the reset step.
extern "C" int step_launch(const double *action, int reset_all) {
    const int n = g.n;
    const size_t NK = (size_t)NUP * NITER * NACT;
    // aabb rows must be 16-byte aligned for double4: the vector's storage is, rows are 32 bytes
    Pool pool{g.obs.data(), g.nv.data(), g.aabb.data(), g.meta.data(), g.nobs.data(), n};
    EnvState st{g.pose.data(), g.cs.data(), g.t.data(), g.accum.data(), g.scene.data(), g.pending.data(), g.gate.data(), g.counters,
                g.traj.data(), g.traj_n.data()};
    const double *t = g.tab.data();
    Tables tb{};
    tb.ray_a = t; tb.ray_b = t + 120; tb.lidar_base = t + 240; tb.mask_base = t + 360; tb.w_lo = t + 480; tb.w_hi = t + 496;
    tb.pmaxk = t + 512; tb.pmax = tb.pmaxk + NK; tb.gpmax = tb.pmax + NUP; tb.maxc = g.maxc;
    hope_out out;
    std::memset(&out, 0, sizeof(out));
    out.pose = g.o_pose.data(); out.target = g.o_target.data(); out.reward = g.o_reward.data(); out.reward_info = g.o_reward_info.data();
    out.status = g.o_status.data(); out.done = g.o_done.data(); out.substeps = g.o_substeps.data(); out.retreated = g.o_retreated.data();
    out.was_reset = g.o_was_reset.data(); out.lidar = g.o_lidar.data(); out.mask = g.o_mask.data(); out.mask_steps = g.o_mask_steps.data();
    out.rs_found = g.o_rs_found.data(); out.rs_nseg = g.o_rs_nseg.data(); out.rs_types = g.o_rs_types.data(); out.rs_lengths = g.o_rs_lengths.data();
    out.rs_L = g.o_rs_L.data(); out.rs_ncand = g.o_rs_ncand.data(); out.rs_ntried = g.o_rs_ntried.data();
    // ---- k_advance ----
    static AdvanceSmem asm_;
    for (int base = 0; base < n; base += 32) {
        const char *err = warp_emu::run([&](int lane) { advance_one(n, base + lane, lane, asm_, pool, st, reset_all ? nullptr : action, g.par, out, reset_all, n); });
        if (err) return fail(err);
    }
    // ---- k_observe ----
    static ObserveSmem osm;
    for (int env = 0; env < n; ++env) {
        const char *err = warp_emu::run([&](int lane) { observe_one(env, lane, osm, pool, st, tb, g.par, out); });
        if (err) return fail(err);
    }
    // ---- k_rs_enumerate: work list in env order ----
    RsScratch rs{};
    int n_items = 0;
    rs.words = g.words.data(); rs.ntry = g.ntry.data(); rs.ncand = g.ncand.data(); rs.item_base = g.item_base.data(); rs.items = g.items.data();
    rs.item_bad = g.item_bad.data(); rs.n_items = &n_items;
    for (int i = 0; i < n; ++i) {
        const int ntry = enumerate_env(i, pool, st, tb, rs, out);
        g.item_base[i] = n_items;
        for (int k = 0; k < ntry; ++k) g.items[n_items++] = (i << 4) | k;
    }
    // ---- k_rs_walk ----
    static std::vector<WordSlot> slots;
    slots.resize(std::max(1, n_items));
    for (int it = 0; it < n_items; ++it) plan_word(slots[it], g.words[(size_t)(g.items[it] >> 4) * MAXW + (g.items[it] & 15)], g.maxc, g.par.rs_step * g.maxc);
    // ---- k_rs_check ----
    unsigned long long votes = 0;
    const unsigned long long div0 = warp_emu::total_work()[0], sc0 = warp_emu::total_work()[1];
    for (int it = 0; it < n_items; ++it) {
        int verdict[32];
        WordSlot &s = slots[it];
        const CheckEnv E = check_env(g.items[it] >> 4);
        const char *err = warp_emu::run([&](int lane) {
            bool bad = false; int chunk_base = 0;
            for (;;) {
#if HOPE_CHK_POOLED
                static CheckSmem cs;
                bad = chunk_is_bad_pooled(s, E, g.par, lane, cs);
#else
                bad = chunk_is_bad(s, E, g.par, lane);
#endif
                if (bad || s.total >= 0) break;
                chunk_base += RS_CHUNK; __syncwarp();
                if (lane == 0) walk_chunk(s, s.len, E.step, chunk_base);
                __syncwarp();
            }
            verdict[lane] = bad ? 1 : 0;
        }, &votes);
        if (err) return fail(err);
        g_check_votes += votes; ++g_check_runs;
        g.item_bad[it] = (uint8_t)verdict[0];
    }
    g_check_div += warp_emu::total_work()[0] - div0; g_check_sincos += warp_emu::total_work()[1] - sc0;
    // ---- k_rs_select ----
    for (int i = 0; i < n; ++i) select_one(i, tb, rs, out);
    return n_items;
}

extern "C" void step_read(double *pose, int32_t *status, double *reward, double *reward_info, double *target, uint8_t *substeps, uint8_t *retreated,
                          double *lidar, double *mask, uint8_t *mask_steps, uint8_t *rs_found, uint8_t *rs_nseg, uint8_t *rs_types, double *rs_lengths,
                          double *rs_L, uint8_t *rs_ncand, uint8_t *rs_ntried) {
    const int n = g.n;
    std::memcpy(pose, g.o_pose.data(), 8 * 3 * n); std::memcpy(status, g.o_status.data(), 4 * n);
    std::memcpy(reward, g.o_reward.data(), 8 * n); std::memcpy(reward_info, g.o_reward_info.data(), 8 * 5 * n);
    std::memcpy(target, g.o_target.data(), 8 * 5 * n); std::memcpy(substeps, g.o_substeps.data(), n); std::memcpy(retreated, g.o_retreated.data(), n);
    std::memcpy(lidar, g.o_lidar.data(), 8 * (size_t)NRAY * n); std::memcpy(mask, g.o_mask.data(), 8 * (size_t)NACT * n);
    std::memcpy(mask_steps, g.o_mask_steps.data(), (size_t)NACT * n);
    std::memcpy(rs_found, g.o_rs_found.data(), n); std::memcpy(rs_nseg, g.o_rs_nseg.data(), n); std::memcpy(rs_types, g.o_rs_types.data(), 5 * n);
    std::memcpy(rs_lengths, g.o_rs_lengths.data(), 8 * 5 * n); std::memcpy(rs_L, g.o_rs_L.data(), 8 * n);
    std::memcpy(rs_ncand, g.o_rs_ncand.data(), n); std::memcpy(rs_ntried, g.o_rs_ntried.data(), n);
}
