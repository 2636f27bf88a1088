// hope_kernels.cu — the ParkingEnv step as sm_100a kernels over N independent scenes.
//
// Kernel            work item          reference code it replaces (src/...)
// k_advance         thread / env       env/env_wrapper.py:37-50, env/vehicle.py:69-96,
//                                      env/car_parking_base.py:153-233, 259-289
// k_observe         warp / env         env/lidar_simulator.py:31-135, model/action_mask.py:166-196,
//                                      env/car_parking_base.py:372-381
// k_rs_enumerate    thread / env       env/reeds_shepp.py:57-449, 540-557; car_parking_base.py:431-444
// k_rs_check        warp / env         env/reeds_shepp.py:452-537, 46-49; car_parking_base.py:452-534
//
// All arithmetic is float64 with FMA contraction off (see hope_device.cuh).  Data layout in HBM is
// documented in DESIGN.md §3.
//
// The per-work-item code of each kernel lives in headers / textual fragments next to this file (advance.cuh + advance_body.inc,
// observe.cuh + observe_body.inc, rs_words.cuh, rs_enumerate.cuh, rs_walk.cuh, rs_check.cuh, rs_select_body.inc, div_pair.cuh,
// hope_types.cuh) so that the host harnesses under tests/ compile the very same source with g++ on a CPU warp emulation
// (tests/warp_emu.h) and replay traces of the unmodified reference through it; this file keeps the kernel shells, the launch
// order and the C ABI.
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hope_b200.h"
#include <sched.h>

#include "hope_device.cuh"
#include "host_wire.h"
#include "scene_gen.h"

namespace hope {

#include "hope_types.cuh"

// k_rs_check, obstacle tests of a round: 0 = the warp votes "any sample hit?" after each obstacle, 1 = after each obstacle
// edge (warp-uniform edge loop).  Same verdicts either way: one bad sample condemns the word.
#ifndef HOPE_CHK_EDGE_EXIT
#define HOPE_CHK_EDGE_EXIT 1
#endif

// Work counters of an instrumented build (-DHOPE_STATS, profiles/tools/kernel_stats.py): where k_rs_check's rounds end
// and how many (quadrant, edge) / (ray, action) items k_observe visits.  The default build contains none of this.
#ifdef HOPE_STATS
__device__ unsigned long long g_stats[64];
#define HOPE_STAT(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define HOPE_STAT(i, v) ((void)0)
#endif

__device__ __forceinline__ double4 ld_aabb(const double4 *p) {  // read-only path, two 128-bit loads
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 lo = __ldg(q), hi = __ldg(q + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// =============================================================================================
// k_advance: one thread per env for the sequential part (pose integration, status, reward); the
// ring-vs-ring collision tests of a warp's 32 envs are pooled and spread over all 32 lanes.
// =============================================================================================
#include "advance.cuh"

// MINB = resident 64-thread blocks per SM the register allocation is sized for.  168 registers (6 blocks) is the
// unconstrained optimum, but 65 536 envs are 1.15 waves of that; capped at 128 registers (8 blocks, a few spills) the
// whole batch is one wave.  Small batches keep the unconstrained build.  64-thread blocks: 1 024 blocks spread over
// 148 SMs within 1 % (512 blocks of 128 leave SMs with 3 or 4).
template <int MINB>
__global__ void __launch_bounds__(ADV_THREADS, MINB) k_advance(int n, Pool pool, EnvState st, const double *__restrict__ action,
                                                       hope_params par, hope_out out, int reset_all, int reset_stride, int raw_action_flag) {
    __shared__ AdvanceSmem smem[ADV_THREADS / 32];
    const bool raw_action = raw_action_flag != 0;
    const int lane = threadIdx.x & 31;
    AdvanceSmem &sm = smem[threadIdx.x >> 5];
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
#include "advance_body.inc"
}

// =============================================================================================
// k_observe: one warp per env.  LiDAR raycast -> action-mask sweep -> target representation.
// =============================================================================================
#include "observe.cuh"

#ifndef HOPE_OBS_MINBLOCKS
#define HOPE_OBS_MINBLOCKS 18  // resident 64-thread blocks per SM the register allocation is sized for (56 registers; shared memory allows 20)
#endif
__global__ void __launch_bounds__(64, HOPE_OBS_MINBLOCKS) k_observe(int n, Pool pool, EnvState st, Tables tb, hope_params par, hope_out out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if (env >= n) return;
    ObserveSmem &sm = reinterpret_cast<ObserveSmem *>(smem_raw)[warp_in_block];
#include "observe_body.inc"
}

// =============================================================================================
// k_rs_enumerate: one thread per env.  46 candidate words -> admitted list -> heap pop order.
// =============================================================================================
#include "render.cuh"

#include "rs_words.cuh"

#include "rs_enumerate.cuh"

__global__ void __launch_bounds__(64) k_rs_enumerate(int n, Pool pool, EnvState st, Tables tb, RsScratch rs, hope_out out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int ntry = i < n ? enumerate_env(i, pool, st, tb, rs, out) : 0;
    // every tried word becomes one work item of k_rs_walk / k_rs_check: warp-level prefix sum, one atomic per warp
    int incl = ntry;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(HOPE_FULL_MASK, incl, o); if (lane >= o) incl += v; }
    const int warp_total = __shfl_sync(HOPE_FULL_MASK, incl, 31);
    int warp_base = 0;
    if (lane == 31 && warp_total) warp_base = atomicAdd(rs.n_items, warp_total);
    warp_base = __shfl_sync(HOPE_FULL_MASK, warp_base, 31);
    if (i < n) {
        const int base = warp_base + incl - ntry;
        rs.item_base[i] = base;
        for (int k = 0; k < ntry; ++k) rs.items[base + k] = (i << 4) | k;
    }
}

// =============================================================================================
// Reeds-Shepp feasibility: k_rs_walk (thread / tried word) -> k_rs_check (warp / tried word)
// -> k_rs_select (thread / env).
//
// Sample each tried word every rs_step metres and test the swept vehicle boxes against map bounds
// and obstacle edges; the first clean word in try order wins (car_parking_base.py:436-450).
//
// The sample positions of a word come from a float accumulation (`pd += d`, reeds_shepp.py:488-492)
// that has to be replayed add by add to reproduce the reference's sample set.  k_rs_walk replays it
// once per word, one word per THREAD (so 32 words advance per warp instruction), and saves a resume
// state every RS_STRIDE samples; k_rs_check gives every word a WARP in which lane j resumes from
// saved state j and evaluates its own RS_STRIDE consecutive samples.  All tried words of an env are
// checked independently (under random actions 93 % of searches fail, i.e. every word is needed
// anyway); k_rs_select then takes the first clean one.
// =============================================================================================
#include "rs_walk.cuh"

__global__ void __launch_bounds__(128) k_rs_walk(RsScratch rs, Tables tb, hope_params par) {
    const int n_items = *rs.n_items;
    const double maxc = tb.maxc, step = par.rs_step * maxc;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        const int code = rs.items[item], env = code >> 4, slot = code & 15;
        plan_word(rs.slots[item], rs.words[(size_t)env * MAXW + slot], maxc, step);
    }
}

#include "rs_check.cuh"
// 1 (shipped) = k_rs_check pools the line-pair tests of a round over the whole warp (rs_check_pooled.cuh): on B200 the step is
// 9 % faster than with each lane walking its obstacle's edges alone (0), all outputs bit-identical
// (profiles/r02_ab_variants.jsonl; the two-words-per-warp, persistent-k_observe and batched-screen variants of round 1
// measured -3 %, +100 % and +20 % and were deleted).
#ifndef HOPE_CHK_POOLED
#define HOPE_CHK_POOLED 1
#endif
#if HOPE_CHK_POOLED
#include "rs_check_pooled.cuh"
#endif

#ifndef HOPE_CHK_WARPS
#define HOPE_CHK_WARPS 4
#endif
#ifndef HOPE_CHK_MINBLOCKS
#define HOPE_CHK_MINBLOCKS 5   // 96 registers, 5 blocks = 20 warps per SM: k_rs_check 7 % faster on B200 than 4 blocks of 116 registers, 6 blocks of 80 no better (profiles/r02_ab_occupancy.jsonl)
#endif
constexpr int CHK_WARPS = HOPE_CHK_WARPS;  // warps per block of k_rs_check

// asynchronous global -> shared copy of one WordSlot by a warp (LDGSTS, 16 bytes per lane per pass)
__device__ __forceinline__ void stage_slot_async(WordSlot *dst, const WordSlot *src, int lane) {
    const char *g = reinterpret_cast<const char *>(src);
    char *sh = reinterpret_cast<char *>(dst);
    for (int q = lane * 16; q < (int)sizeof(WordSlot); q += 32 * 16) __pipeline_memcpy_async(sh + q, g + q, 16);
    __pipeline_commit();
}

__global__ void __launch_bounds__(CHK_WARPS * 32, HOPE_CHK_MINBLOCKS) k_rs_check(Pool pool, EnvState st, Tables tb, RsScratch rs, hope_params par) {
    __shared__ WordSlot smem[CHK_WARPS][2];  // double buffer: the next word's plan streams in while this one is sampled
#if HOPE_CHK_POOLED
    __shared__ CheckSmem csm[CHK_WARPS];
#endif
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = *rs.n_items;
    const int warps_total = gridDim.x * CHK_WARPS;
    int item = blockIdx.x * CHK_WARPS + warp_in_block, buf = 0;
    if (item < n_items) stage_slot_async(&smem[warp_in_block][0], rs.slots + item, lane);
    for (; item < n_items; item += warps_total, buf ^= 1) {
        const int env = rs.items[item] >> 4;
        const int next = item + warps_total;
        if (next < n_items) stage_slot_async(&smem[warp_in_block][buf ^ 1], rs.slots + next, lane);
        const int sid = st.scene[env];
        const double *meta = pool.meta + (size_t)sid * META;
        CheckEnv E;
        E.q0x = st.pose[3 * env]; E.q0y = st.pose[3 * env + 1]; E.q0h = st.pose[3 * env + 2];
        E.cg = st.cs[2 * env]; E.sg = -st.cs[2 * env + 1];  // cos(-h), sin(-h)  (reeds_shepp.py:47-48)
        E.xmin = meta[M_BOUNDS]; E.xmax = meta[M_BOUNDS + 1]; E.ymin = meta[M_BOUNDS + 2]; E.ymax = meta[M_BOUNDS + 3];
        E.maxc = tb.maxc; E.step = par.rs_step * tb.maxc;
        E.nobs = pool.nobs[sid];
        E.aabb = reinterpret_cast<const double4 *>(pool.aabb) + (size_t)sid * MAXO;
        E.verts = reinterpret_cast<const double2 *>(pool.obs) + (size_t)sid * MAXE;
        E.nvp = pool.nv + (size_t)sid * MAXO;
        if (next < n_items) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);  // this word's plan has landed
        __syncwarp();
        WordSlot &s = smem[warp_in_block][buf];
        bool bad = false;
        int chunk_base = 0;
        for (;;) {
#if HOPE_CHK_POOLED
            bad = chunk_is_bad_pooled(s, E, par, lane, csm[warp_in_block]);
#else
            bad = chunk_is_bad(s, E, par, lane);
#endif
            if (bad || s.total >= 0) break;
            chunk_base += RS_CHUNK;  // a word longer than one chunk (rare): lane 0 walks on
            __syncwarp();
            if (lane == 0) walk_chunk(s, s.len, E.step, chunk_base);
            __syncwarp();
        }
        if (lane == 0) rs.item_bad[item] = bad ? 1 : 0;
#ifdef HOPE_STATS
        if (lane == 0) { HOPE_STAT(0, 1); if (bad) HOPE_STAT(1, 1); if (chunk_base) HOPE_STAT(10, 1); if (s.total >= 0) { HOPE_STAT(11, s.total); HOPE_STAT(35, 1); } }
#endif
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) k_rs_select(int n, Tables tb, RsScratch rs, hope_out out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#include "rs_select_body.inc"
}

// =============================================================================================
// k_planner: one thread per env.  RsPlanner.set_rs_path / get_action evaluated lazily.
// =============================================================================================
struct PlanState {
    double *rem;       // [N][5] remaining signed length of each segment in policy units (length / step_ratio)
    uint8_t *types;    // [N][5]
    uint8_t *big;      // [N]    bit k: segment k started with |a| > 1 (its remainder rule differs, parking_agent.py:26-37)
    uint8_t *n;        // [N]    segments
    uint8_t *seg;      // [N]    current segment
    uint8_t *active;   // [N]    planner.route is not None
};

// next action of the filtered list, or false when the list is exhausted.  `commit` = pop it.
__device__ __forceinline__ bool plan_next(double *rem, const uint8_t *types, unsigned big, int n, int &seg, bool commit, double &steer, double &speed) {
    double local[HOPE_RS_MAX_SEG];
    if (!commit) for (int k = 0; k < HOPE_RS_MAX_SEG; ++k) local[k] = rem[k];
    double *r = commit ? rem : local;
    while (seg < n) {
        const double a = r[seg];
        const int t = types[seg];
        steer = t == HOPE_RS_L ? 1.0 : (t == HOPE_RS_R ? -1.0 : 0.0);  // action_type {'L':1,'S':0,'R':-1}
        if (!((big >> seg) & 1)) {  // |a| <= 1 from the start: one action if 1e-3 < |a| < 1, else nothing (|a| == 1 is dropped, sic)
            ++seg;
            if (fabs(a) < 1.0 && fabs(a) > 1e-3) { speed = a; return true; }
        } else if (a > 1.0) { r[seg] = a - 1.0; speed = 1.0; return true; }
        else if (a < -1.0) { r[seg] = a + 1.0; speed = -1.0; return true; }
        else {  // remainder of a long segment
            ++seg;
            if (fabs(a) > 1e-3) { speed = a; return true; }
        }
    }
    return false;
}

__global__ void __launch_bounds__(128) k_planner(int n, PlanState ps, const double *__restrict__ policy_action, hope_out last,
                                                 double *__restrict__ action_out, uint8_t *__restrict__ executing, double step_ratio) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool active = ps.active[i] != 0;
    const bool ended = (last.done && last.done[i]) || (last.was_reset && last.was_reset[i]);
    if (ended) active = false;  // ParkingAgent.reset at the start of the next episode (parking_agent.py:60-62)
    else if (!active && last.rs_found && last.rs_found[i]) {  // set_planner_path: only when no route is being executed (:64-68)
        const int nseg = last.rs_nseg[i];
        unsigned big = 0;
        for (int k = 0; k < HOPE_RS_MAX_SEG; ++k) {
            const double a = k < nseg ? last.rs_lengths[5 * i + k] / step_ratio : 0.0;  // :16-17
            ps.rem[5 * i + k] = a;
            ps.types[5 * i + k] = k < nseg ? last.rs_types[5 * i + k] : HOPE_RS_NONE;
            if (fabs(a) > 1.0) big |= 1u << k;
        }
        ps.big[i] = (uint8_t)big; ps.n[i] = (uint8_t)nseg; ps.seg[i] = 0;
        active = true;
    }
    double steer = policy_action[2 * i], speed = policy_action[2 * i + 1];
    bool from_plan = false;
    if (active) {
        int seg = ps.seg[i];
        double ps_steer, ps_speed;
        if (plan_next(ps.rem + 5 * i, ps.types + 5 * i, ps.big[i], ps.n[i], seg, true, ps_steer, ps_speed)) {
            steer = ps_steer; speed = ps_speed; from_plan = true;
            int peek = seg;
            double d0, d1;
            // RsPlanner.get_action: the route is dropped as soon as its last action is popped (:42-46)
            if (!plan_next(ps.rem + 5 * i, ps.types + 5 * i, ps.big[i], ps.n[i], peek, false, d0, d1)) active = false;
        } else active = false;  // empty filtered list (the reference would raise IndexError on pop)
        ps.seg[i] = (uint8_t)seg;
    }
    ps.active[i] = active ? 1 : 0;
    action_out[2 * i] = steer; action_out[2 * i + 1] = speed;
    if (executing) executing[i] = from_plan ? 1 : 0;
}

// envs that finished last step -> dense list of their pool slots (regen_on_reset); bumps the slot's episode counter
__global__ void __launch_bounds__(128) k_list_pending(int n, const uint8_t *__restrict__ pending, const int *__restrict__ scene,
                                                      unsigned *__restrict__ episode, int *__restrict__ slots, int *__restrict__ count,
                                                      unsigned long long *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool p = i < n && pending[i] != 0;
    const unsigned m = __ballot_sync(HOPE_FULL_MASK, p);
    int base = 0;
    if (lane == 0 && m) { base = atomicAdd(count, __popc(m)); atomicAdd(counters + 6, (unsigned long long)__popc(m)); }
    base = __shfl_sync(HOPE_FULL_MASK, base, 0);
    if (p) {
        const int slot = scene[i];
        episode[slot] += 1;
        slots[base + __popc(m & ((1u << lane) - 1))] = slot;
    }
}
// pad the rest of the list with -1 so the generator kernel can be launched over all n entries
__global__ void k_pad_slots(int n, int *__restrict__ slots, const int *__restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i >= *count) slots[i] = -1;
}

// small helpers -------------------------------------------------------------------------------
__global__ void k_table_reduce(const double *__restrict__ dist_star, double *__restrict__ pmaxk, double *__restrict__ pmax) {
    // one block per upsampled ray: pmaxk[rho][k][j] = max_{k'<=k} dist_star[rho][j][k'], pmax[rho] = max_j pmaxk[rho][9][j]
    int rho = blockIdx.x, j = threadIdx.x;
    __shared__ double sh[64];
    double m = -1.0;
    if (j < NACT) {
        const double *row = dist_star + ((size_t)rho * NACT + j) * NITER;
        for (int k = 0; k < NITER; ++k) {
            m = k == 0 ? row[0] : dmax(m, row[k]);
            pmaxk[((size_t)rho * NITER + k) * NACT + j] = m;
        }
    }
    sh[j] = m;
    __syncthreads();
    if (j == 0) {
        double mm = sh[0];
        for (int q = 1; q < NACT; ++q) mm = dmax(mm, sh[q]);
        pmax[rho] = mm;
    }
}

// float64 FMA throughput probe: 8 independent chains per thread so the pipe, not the latency, is measured
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void k_bump_seq(unsigned long long *seq) { *seq += 1ull; }

// k_pack_lidar: one warp per env, behind k_observe in the host API.  A beam that hits nothing within range reads exactly
// lidar_range - lidar_base[ray] (observe_body.inc: clip to the range, subtract the vehicle's extent; lidar_simulator.py:46,134),
// 43 % of the beams of a random-action step.  Only the values whose BITS differ from that per-ray constant are kept: bits[env]
// flags them (bit j of word q = ray 32 q + j), off[env] is where the env's kept values start in its sub-range's region of
// `packed` (one atomic per env on the sub-range's counter, so the order of envs in `packed` varies from run to run while the
// expansion on the host does not).
// Streaming kernel: reads 960 B per env (still in L2 behind k_observe), writes 20 B + 8 B per kept beam.
static_assert(NRAY <= 128, "k_pack_lidar: four 32-bit flag words per env");
__global__ void __launch_bounds__(128) k_pack_lidar(int n, const double *__restrict__ lidar, const double *__restrict__ lidar_base, double range,
                                                    uint32_t *__restrict__ bits, uint32_t *__restrict__ off, double *__restrict__ packed,
                                                    unsigned *__restrict__ count, int sub) {
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (env >= n) return;
    const int g = env / sub;          // sub-range of the launch's env range: own counter, own region of `packed`
    count += g;
    packed += (size_t)g * sub * NRAY;
    const double *row = lidar + (size_t)env * NRAY;
    double v[4];
    unsigned m[4];
    int total = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = q * 32 + lane;
        const bool in = j < NRAY;
        v[q] = in ? row[j] : 0.0;
        const double nohit = in ? range - __ldg(lidar_base + j) : 0.0;
        m[q] = __ballot_sync(HOPE_FULL_MASK, in && __double_as_longlong(v[q]) != __double_as_longlong(nohit));
        total += __popc(m[q]);
    }
    unsigned base = 0;
    if (lane == 0 && total) base = atomicAdd(count, (unsigned)total);
    base = __shfl_sync(HOPE_FULL_MASK, base, 0);
    if (lane == 0) {
        *reinterpret_cast<uint4 *>(bits + 4 * (size_t)env) = make_uint4(m[0], m[1], m[2], m[3]);
        off[env] = base;
    }
    unsigned pos = base;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if ((m[q] >> lane) & 1) packed[pos + __popc(m[q] & ((1u << lane) - 1))] = v[q];
        pos += __popc(m[q]);
    }
}

__global__ void k_table_group(const double *__restrict__ pmax, double *__restrict__ gpmax) {
    int q = threadIdx.x;
    if (q >= NRAY) return;
    double m = pmax[q * 10];
    for (int r = 1; r < 10; ++r) m = fmax(m, pmax[q * 10 + r]);
    gpmax[q] = m;
}

}  // namespace hope

// =============================================================================================
// Host side: context and C ABI
// =============================================================================================
using namespace hope;

// A few persistent host threads for the one piece of host-side work in hope_step_host (expanding the uint8
// action-mask steps into the float64 mask while the remaining copies are still travelling).
struct HostPool {
    // Jobs are queued by the stepping thread and run in order; every worker runs its own share (part, nparts) of each job
    // and moves on without waiting for the others (the jobs of one step write disjoint memory).  A worker that runs out of
    // jobs polls for spin_us microseconds before it sleeps on the condition variable: the jobs of a step arrive 0.1 - 0.3 ms
    // apart, and a futex wake-up costs about as much as a job.
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::vector<std::function<void(int, int)>> jobs;  // of the current step; cleared by wait_all
    std::vector<size_t> progress;                      // per worker: jobs finished
    std::atomic<size_t> submitted{0};                  // jobs queued since the pool started (never reset), readable without the lock
    std::atomic<bool> stop{false};
    int spin_us = 200;
    void start(int n) {
        progress.assign(n, 0);
        for (int w = 0; w < n; ++w)
            workers.emplace_back([this, w, n] {
                size_t mine = 0;  // jobs this worker has finished since the pool started
                for (;;) {
                    if (spin_us > 0) {
                        const auto t0 = std::chrono::steady_clock::now();
                        for (unsigned k = 0; submitted.load(std::memory_order_acquire) <= mine && !stop.load(std::memory_order_relaxed); ++k) {
#if defined(__x86_64__)
                            __builtin_ia32_pause();
#endif
                            if ((k & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us)) break;
                        }
                    }
                    std::function<void(int, int)> f;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&] { return stop.load() || progress[w] < jobs.size(); });
                        if (stop.load()) return;
                        f = jobs[progress[w]];
                    }
                    f(w, n);
                    ++mine;
                    std::lock_guard<std::mutex> lk(m);
                    if (++progress[w] == jobs.size()) done_cv.notify_one();
                }
            });
    }
    void submit(std::function<void(int, int)> f) {
        { std::lock_guard<std::mutex> lk(m); jobs.push_back(std::move(f)); submitted.fetch_add(1, std::memory_order_release); }
        cv.notify_all();
    }
    void wait_all() {
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [&] { for (size_t p : progress) if (p < jobs.size()) return false; return true; });
        jobs.clear();
        for (size_t &p : progress) p = 0;
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m); stop.store(true); }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
};

struct hope_ctx {
    int device = 0, n = 0, pool = 0;
    hope_params par;
    // pool
    double *d_obs = nullptr, *d_aabb = nullptr, *d_meta = nullptr;
    uint8_t *d_nv = nullptr;
    int *d_nobs = nullptr;
    // tables
    double *d_tab = nullptr;  // ray_a ray_b lidar_base mask_base w_lo w_hi | dist_star | pend | pmax
    bool have_tables = false, have_scenes = false, have_reset = false;
    double maxc = 0.0;
    // state
    double *d_pose = nullptr, *d_accum = nullptr, *d_cs = nullptr;
    int *d_t = nullptr, *d_scene = nullptr;
    uint8_t *d_pending = nullptr, *d_gate = nullptr;
    unsigned long long *d_counters = nullptr;
    // RS scratch
    RsWord *d_words = nullptr;
    uint8_t *d_ntry = nullptr, *d_ncand = nullptr, *d_item_bad = nullptr;
    int *d_item_base = nullptr, *d_items = nullptr, *d_n_items = nullptr;
    WordSlot *d_slots = nullptr;
    int *d_regen_slots = nullptr, *d_regen_count = nullptr, *d_gen_status = nullptr;
    unsigned *d_episode = nullptr;
    double *d_traj = nullptr;
    int *d_traj_n = nullptr;
    render::Palette palette;
    render::Camera *d_cams = nullptr;
    uint8_t *d_screen = nullptr;      // [N][500][125]: the static part of every env's screen, 2 bits per pixel (k_render_static), allocated with the first image
    uint2 *d_screen_key = nullptr;
    uint8_t *d_spanrec = nullptr;     // [N][render::SPANREC]: what each box of the trajectory ring buffer paints on its screen rows, with the pose it was computed from
    int *d_repaint = nullptr, *d_repaint_n = nullptr;  // [N] each: envs whose static screen is stale (k_render_camera -> k_render_static), per env range: list at [lo ..), its length at [lo]
    int render_traj_len = render::TRAJ;  // hope_set_render_traj: configs.py:86 TRAJ_RENDER_LEN (0: RENDER_TRAJ off)
    int render_force_lattice = 0;
    size_t render_pad = 0;            // HOPE_B200_RENDER_PAD_KB: extra dynamic shared memory per k_render CTA (fewer resident CTAs, room for the step kernels beside them)     // HOPE_B200_RENDER_LATTICE=1 (tests): k_render resolves the dynamic layer per lattice sample even when its window fits    // [N]: (pool slot, its regeneration count) the cached screen was painted for; all ones = none
    double *d_plan_rem = nullptr;
    uint8_t *d_plan_u8 = nullptr;  // types[N][5] | big[N] | n[N] | seg[N] | active[N]
    int sm_count = 148, walk_blocks = 148 * 4, check_blocks = 148 * 4, host_check_blocks = 148 * 4;
    // host-API staging
    double *d_action = nullptr;
    void *d_stage = nullptr, *d_stage_img = nullptr;
    size_t stage_bytes = 0;
    hope_out stage_out;
    cudaStream_t own_stream = nullptr;
    // a "lane" = the stream pair one env range is stepped on: k_observe goes to `aux`, concurrently with the
    // Reeds-Shepp kernels on `main`.  Lane 0's main stream is replaced by the caller's in the device API; the
    // host API pipelines env ranges over all lanes so one range's D2H copies hide under the next one's kernels.
    static constexpr int MAX_LANES = 4;
    struct Lane { cudaStream_t main = nullptr, aux = nullptr; cudaEvent_t ev_advanced = nullptr, ev_observed = nullptr; } lanes[MAX_LANES];
    int host_chunks = 2;   // measured on B200 at 65 536 envs with the narrow mask format: 2 -> 2.18 ms, 3 -> 2.21, 4 -> 2.35 per
                           // host step (with the float64 mask copied: 2 -> 2.60, 3 -> 2.53, 4 -> 2.63, 8 -> 3.30)
    bool host_rs_after_observe = false;  // HOPE_B200_HOST_RS_AFTER=1: start the Reeds-Shepp kernels behind the last k_observe range (better when the copies, not the kernels, bound the step)
    bool in_host_step = false;
    bool render_after_rs = false;
    bool image_ok = true;  // false: the vehicle box is too large for k_render's per-box span table
    // Narrow wire format of the host API: the float64 action mask (336 B per env, 23 % of the bytes a step returns) is
    // a function of its 42 uint8 step counts (k_observe: steps / 10, or 0.01 everywhere when all are 0), so the
    // counts travel and a few host threads expand them into the caller's float64 buffer while the lidar copies are
    // still in flight.  HOPE_B200_HOST_MASK_EXPAND=0 copies the float64 mask instead.
    bool host_mask_expand = true, expanding = false;
    uint8_t *h_mask_steps = nullptr;   // pinned [N][42]
    // Same idea for the lidar (960 B per env, 84 % of the bytes): beams that hit nothing read a per-ray constant, so
    // k_pack_lidar keeps the others and the host threads rebuild the float64 rows (host_wire.cpp).  The kept values of an env
    // range are copied by a cudaMemcpyAsync the stepping thread issues as soon as the range's count has landed (its size is
    // only known then, so it cannot be part of the replayed graph).  HOPE_B200_HOST_LIDAR_PACK=0 copies the float64 rows.
    bool host_lidar_pack = true, packing = false;
    uint32_t *d_lbits = nullptr, *d_loff = nullptr, *h_lbits = nullptr, *h_loff = nullptr;   // [N][4], [N]
    double *d_lpacked = nullptr, *h_lpacked = nullptr;                                        // [N][120] worst case; range c uses [lo_c * 120, ...)
    unsigned *d_lcount = nullptr, *h_lcount = nullptr;                                        // [MAX_CHUNK_EVENTS] kept values per range
    cudaStream_t s_pack = nullptr;
    cudaEvent_t ev_pack[64] = {};
    // an env range's kept values are packed, copied and expanded in sub-ranges of pk_sub envs (own counter, own copy, own
    // expansion job), so the host work left when the last copy lands is one sub-range, not one range
    int pk_sub = 2048, pk_total = 0;
    int pk_lo[64] = {}, pk_hi[64] = {}, pk_first[65] = {};   // sub-range g covers envs [pk_lo, pk_hi); range c owns sub-ranges [pk_first[c], pk_first[c+1])
    // Packing trades PCIe bytes (37 instead of 63 MB of lidar per 65 536-env step) for host work (the rows are rebuilt by CPU
    // stores).  With one rank per box and a dozen host threads the trade wins; with 8 ranks sharing 32 CPUs the expansion
    // becomes the bottleneck (B200 x8: 9.8 ms per step all packed, 8.0 ms none packed).  So only pack_frac of the sub-ranges
    // travel packed, spread evenly; the others' float64 rows are copied straight into the caller's buffer.  Both kinds of copy
    // are issued by the stepping thread, outside the replayed graph, so the split can change from step to step.  Default: all
    // packed for one rank per box, 3/4 for two, none for more.
    double pack_frac = 1.0;
    int ranks_per_box = 1;
    bool pk_packed[64] = {};
    double h_nohit[HOPE_N_LIDAR] = {};   // lidar_range - lidar_base[ray], the same float64 subtraction k_observe performs
    int wire_force_portable = 0;
    bool host_noexpand = false;  // HOPE_B200_HOST_NOEXPAND=1: timing experiments only (the host arrays are not rebuilt)
    // HOPE_B200_HOST_TRACE=k: the k-th hope_step_host call (and every 50th after it) is enqueued directly instead of replayed
    // from the graph, with timing events between its parts, and its device and host timelines are printed to stderr as JSON
    int host_trace = 0;
    bool tracing = false;
    unsigned long long host_steps = 0;
    std::vector<std::pair<std::string, cudaEvent_t>> trace_ev;
    std::vector<std::pair<std::string, double>> trace_host;
    std::chrono::steady_clock::time_point trace_t0;
    unsigned long long io_h2d = 0, io_d2h = 0, io_d2h_static = 0;   // bytes of the last hope_step_host (static = the part enqueued / captured)
    HostPool *host_pool = nullptr;
    int host_threads = 0;
    int hm_chunks = 0;
    int ch_lo[65] = {};      // env range c of the pipelined host step = [ch_lo[c], ch_lo[c + 1])
    std::string host_split;  // HOPE_B200_HOST_SPLIT="50,30,20": the ranges as percentages of the envs instead of host_chunks equal ones
    // "range c's step counts have landed": a device sequence number, bumped once per host step, is copied into
    // pinned host memory right behind each range's step counts (same stream, so in order); the host spins on it.
    // (An event recorded inside a replayed graph cannot be used for this: until the node runs it still reports the
    // previous launch's completed record.)
    unsigned long long *d_seq = nullptr, host_seq = 0;
    volatile unsigned long long *h_seq = nullptr;  // pinned [MAX_CHUNK_EVENTS]
    int host_debug = 0;  // HOPE_B200_HOST_DEBUG: 1 = enqueue no copies, 2 = enqueue no kernels (timing experiments only)
    static constexpr int MAX_CHUNK_EVENTS = 64;
    cudaEvent_t ev_chunk[MAX_CHUNK_EVENTS] = {}, ev_adv[MAX_CHUNK_EVENTS] = {};
    bool host_split_advance = true;
    int device_chunks = 1;  // hope_step: env ranges stepped on separate lanes so one range's latency-bound kernels
                            // (advance, enumerate, walk) run under another range's issue-bound ones (observe, check)
    // hope_step_host replays a captured CUDA graph of the whole pipelined step while the caller keeps passing the
    // same buffers (one launch instead of ~120 driver calls per step)
    bool host_graph_enabled = true;
    // Optional (HOPE_B200_ZERO_COPY=1): outputs whose host buffer is pinned + mapped (cudaHostAlloc / cudaHostRegister,
    // e.g. torch pin_memory) are written by the kernels directly over PCIe instead of staged in HBM and copied:
    // bit k = field k of kOutFields.  3.3 ms vs 2.7 ms per 65 536-env step on B200, so staged copies stay the default.
    bool zero_copy_enabled = false;  // measured slower than staged copies on B200 (small PCIe write TLPs): opt-in
    unsigned zero_copy_mask = 0;
    hope_out step_out;   // what the kernels of the current host step write to (stage_out with mapped pointers patched in)
    cudaGraphExec_t host_graph = nullptr;
    const double *hg_action = nullptr;
    hope_out hg_out;
    unsigned hg_stages = 0;
    unsigned long long hg_launches = 0;
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    // The device API (hope_reset / hope_step / hope_planner_actions) is asynchronous on the CALLER's stream, the host API
    // runs on the context's own non-blocking streams: the last device-API call leaves an event here and the host API
    // waits for it, so the two can be mixed without an explicit synchronisation in between.
    cudaEvent_t ev_dev = nullptr;
    bool dev_pending = false;
    cudaEvent_t ev_obs_done = nullptr;  // recorded behind k_observe [+ k_render] of the last hope_step / hope_reset over the whole batch
    bool obs_done_valid = false;
    unsigned long long launches = 0;
    bool profile = false;
    bool profile_serial = false;  // hope_profile_enable(ctx, 2): no k_observe / Reeds-Shepp overlap, so each kernel is timed alone
    std::vector<cudaEvent_t> prof_events[8];  // begin/end pairs per kernel
    std::string last_error;
};

namespace {

int fail(hope_ctx *c, cudaError_t e, const char *what) {
    if (c) c->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return HOPE_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail(ctx, e_, #call);        \
    } while (0)

Pool make_pool(const hope_ctx *c) { return Pool{c->d_obs, c->d_nv, c->d_aabb, c->d_meta, c->d_nobs, c->pool}; }
Tables make_tables(const hope_ctx *c) {
    const double *t = c->d_tab;
    Tables tb;
    tb.ray_a = t; tb.ray_b = t + 120; tb.lidar_base = t + 240; tb.mask_base = t + 360; tb.w_lo = t + 480; tb.w_hi = t + 496;
    tb.dist_star = t + 512; tb.pmaxk = tb.dist_star + (size_t)NUP * NACT * NITER; tb.pmax = tb.pmaxk + (size_t)NUP * NACT * NITER; tb.gpmax = tb.pmax + NUP;
    tb.maxc = c->maxc;
    return tb;
}
constexpr int OBS_THREADS = 64, ENUM_THREADS = 64;

void prof_mark(hope_ctx *ctx, int which, cudaStream_t s) {
    if (!ctx->profile) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    ctx->prof_events[which].push_back(e);
}

void tmark(hope_ctx *ctx, const char *what, int idx, cudaStream_t s) {  // device timeline of a traced host step
    if (!ctx->tracing) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    ctx->trace_ev.emplace_back(idx >= 0 ? std::string(what) + "[" + std::to_string(idx) + "]" : std::string(what), e);
}
void hmark(hope_ctx *ctx, const char *what, int idx) {  // host timeline of a traced host step
    if (!ctx->tracing) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ctx->trace_t0).count();
    ctx->trace_host.emplace_back(idx >= 0 ? std::string(what) + "[" + std::to_string(idx) + "]" : std::string(what), ms);
}
void trace_print(hope_ctx *ctx) {
    if (!ctx->tracing) return;
    std::string out = "{\"host_step\": " + std::to_string(ctx->host_steps) + ", \"device_ms\": {";
    for (size_t k = 0; k < ctx->trace_ev.size(); ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->trace_ev[0].second, ctx->trace_ev[k].second);
        char buf[64]; snprintf(buf, sizeof(buf), "%.4f", ms);
        out += (k ? ", \"" : "\"") + ctx->trace_ev[k].first + "\": " + buf;
    }
    out += "}, \"host_ms\": {";
    for (size_t k = 0; k < ctx->trace_host.size(); ++k) {
        char buf[64]; snprintf(buf, sizeof(buf), "%.4f", ctx->trace_host[k].second);
        out += (k ? ", \"" : "\"") + ctx->trace_host[k].first + "\": " + buf;
    }
    out += "}}\n";
    fputs(out.c_str(), stderr);
    for (auto &e : ctx->trace_ev) cudaEventDestroy(e.second);
    ctx->trace_ev.clear(); ctx->trace_host.clear();
    ctx->tracing = false;
}

// Kernel order of one step.  k_observe and the Reeds-Shepp pair both depend only on k_advance, so when
// both stages are requested k_observe goes to the context's auxiliary stream and overlaps the RS kernels;
// the caller's stream waits for it before hope_step returns control of the stream.  With `early_out`
// (host API) the observation buffers are copied to the host right behind k_observe, under the RS kernels.
enum { BY_ADVANCE = 1, BY_OBSERVE = 2, BY_RS = 4, BY_ANY = 7 };  // which kernel produces an output array
struct OutField { size_t offset; size_t elem; int per_env; int by; };
#define OF(member, type, per) OutField{offsetof(hope_out, member), sizeof(type), per, BY_ADVANCE}
#define OFO(member, type, per) OutField{offsetof(hope_out, member), sizeof(type), per, BY_OBSERVE}
#define OFR(member, type, per) OutField{offsetof(hope_out, member), sizeof(type), per, BY_RS}
const OutField kOutFields[] = {
    OF(pose, double, 3), OFO(lidar, double, NRAY), OFO(mask, double, NACT), OFO(mask_steps, uint8_t, NACT),
    OF(target, double, 5), OF(reward, double, 1), OF(reward_info, double, 5), OF(status, int32_t, 1),
    OF(done, uint8_t, 1), OF(substeps, uint8_t, 1), OF(retreated, uint8_t, 1), OF(was_reset, uint8_t, 1),
    OFR(rs_found, uint8_t, 1), OFR(rs_nseg, uint8_t, 1), OFR(rs_types, uint8_t, 5), OFR(rs_lengths, double, 5),
    OFR(rs_L, double, 1), OFR(rs_ncand, uint8_t, 1), OFR(rs_ntried, uint8_t, 1), OFO(img, uint8_t, HOPE_IMG_C * HOPE_IMG_HW * HOPE_IMG_HW)};
constexpr int kNumOutFields = sizeof(kOutFields) / sizeof(kOutFields[0]);

void *&field_ptr(hope_out &o, const OutField &f) { return *reinterpret_cast<void **>(reinterpret_cast<char *>(&o) + f.offset); }
void *field_ptr_c(const hope_out &o, const OutField &f) { return *reinterpret_cast<void *const *>(reinterpret_cast<const char *>(&o) + f.offset); }

hope_out offset_out(const hope_out &o, size_t lo) {  // the same arrays, starting at env `lo`
    hope_out r = o;
    for (int k = 0; k < kNumOutFields; ++k) {
        void *p = field_ptr_c(o, kOutFields[k]);
        if (p) field_ptr(r, kOutFields[k]) = static_cast<char *>(p) + lo * kOutFields[k].per_env * kOutFields[k].elem;
    }
    return r;
}

// by: the arrays of which producers to copy (BY_* bits); envs [lo, lo+cnt)
int copy_fields(hope_ctx *ctx, const hope_host_out *h_out, int by, cudaStream_t s, size_t lo, size_t cnt) {
    if (ctx->host_debug == 1 && ctx->in_host_step) return HOPE_OK;
    for (int k = 0; k < kNumOutFields; ++k) {
        void *dst = field_ptr_c(*h_out, kOutFields[k]);
        if (!dst || !(kOutFields[k].by & by)) continue;
        if ((ctx->zero_copy_mask >> k) & 1) continue;  // the kernels wrote this array straight into the caller's mapped buffer
        if (ctx->expanding && ctx->in_host_step && kOutFields[k].offset == offsetof(hope_out, mask)) continue;  // rebuilt on the host
        if (ctx->packing && ctx->in_host_step && kOutFields[k].offset == offsetof(hope_out, lidar)) continue;   // rebuilt on the host
        const size_t row = kOutFields[k].elem * kOutFields[k].per_env;
        if (ctx->in_host_step) ctx->io_d2h_static += row * cnt;
        CK(cudaMemcpyAsync(static_cast<char *>(dst) + lo * row, static_cast<const char *>(field_ptr_c(ctx->stage_out, kOutFields[k])) + lo * row,
                           row * cnt, cudaMemcpyDeviceToHost, s));
    }
    return HOPE_OK;
}

// Kernel order of one step over envs [lo, lo+cnt).  k_observe and the Reeds-Shepp kernels both depend only on
// k_advance, so when both stages are requested k_observe goes to the lane's auxiliary stream and overlaps the RS
// kernels; the main stream waits for it at the end.  With `early_out` (host API) the observation buffers are
// copied to the host right behind k_observe, under the RS kernels.
int launch_range(hope_ctx *ctx, const double *d_action, const hope_out &out_all, unsigned stages, int reset_all, cudaStream_t s,
                 int lane_id, int chunk_id, int lo, int cnt, const hope_host_out *early_out = nullptr, bool do_advance = true) {
    const int n = cnt;
    if (ctx->host_debug == 2 && ctx->in_host_step) return HOPE_OK;
    hope_ctx::Lane &lane = ctx->lanes[lane_id];
    Pool pool = make_pool(ctx);
    Tables tb = make_tables(ctx);
    EnvState st{ctx->d_pose + 3 * (size_t)lo, ctx->d_cs + 2 * (size_t)lo, ctx->d_t + lo, ctx->d_accum + lo, ctx->d_scene + lo,
                ctx->d_pending + lo, ctx->d_gate + lo, ctx->d_counters, ctx->d_traj + 80 * (size_t)lo, ctx->d_traj_n + lo};
    const hope_out out = offset_out(out_all, lo);
    const double *act = d_action ? d_action + 2 * (size_t)lo : nullptr;
    const bool regen = ctx->par.regen_on_reset && ctx->par.auto_reset && !reset_all;
    if (regen && do_advance) {  // fresh scenes for the envs that finished last step, generated in place on the device
        int *slots = ctx->d_regen_slots + lo, *count = ctx->d_n_items + 32 + chunk_id % 32;
        CK(cudaMemsetAsync(count, 0, sizeof(int), s));
        k_list_pending<<<(n + 127) / 128, 128, 0, s>>>(n, st.pending, st.scene, ctx->d_episode, slots, count, ctx->d_counters);
        k_pad_slots<<<(n + 127) / 128, 128, 0, s>>>(n, slots, count);
        if (hope_scene::launch_generate(n, 0, ctx->par.regen_level, ctx->par.regen_seed, slots, ctx->d_episode, ctx->par, ctx->d_obs, ctx->d_nv,
                                        ctx->d_aabb, ctx->d_meta, ctx->d_nobs, ctx->d_gen_status, s) != HOPE_OK) return HOPE_ERR_CUDA;
        ctx->launches += 3;
    }
    if (do_advance) {
        prof_mark(ctx, 0, s);
        const int adv_blocks = (n + ADV_THREADS - 1) / ADV_THREADS;
        const int raw = (stages & HOPE_STAGE_RAW_ACTION) ? 1 : 0;
        if (adv_blocks > 6 * ctx->sm_count)  // more than one wave at 6 blocks per SM
            k_advance<8><<<adv_blocks, ADV_THREADS, 0, s>>>(n, pool, st, act, ctx->par, out, reset_all, regen ? 0 : ctx->n, raw);
        else
            k_advance<6><<<adv_blocks, ADV_THREADS, 0, s>>>(n, pool, st, act, ctx->par, out, reset_all, regen ? 0 : ctx->n, raw);
        prof_mark(ctx, 0, s);
        ctx->launches++;
    }
    const bool image = (stages & HOPE_STAGE_IMAGE) && out.img;
    if (image && !ctx->image_ok) {
        ctx->last_error = "image observation: the vehicle box spans more screen rows than k_render's span table holds (render::DROWS)";
        return HOPE_ERR_CAPACITY;
    }
    // optional (HOPE_B200_RENDER_AFTER_RS=1): k_render behind the Reeds-Shepp kernels on the main stream instead of
    // next to them on the auxiliary one; measured equal within 2 % (11.0 vs 10.8 ms per 65 536-env step), so off
    const bool image_last = image && (stages & HOPE_STAGE_RS) && !ctx->in_host_step && ctx->render_after_rs;
    const bool side = (stages & HOPE_STAGE_OBSERVE) || (image && !image_last);  // work that only depends on k_advance, besides RS
    const bool fork = side && (stages & HOPE_STAGE_RS) && !(ctx->profile && ctx->profile_serial);
    cudaStream_t so = fork ? lane.aux : s;
    if (side) {
        if (fork) {
            CK(cudaEventRecord(lane.ev_advanced, s));
            CK(cudaStreamWaitEvent(so, lane.ev_advanced, 0));
        }
        if (stages & HOPE_STAGE_OBSERVE) {
            const int wpb = OBS_THREADS / 32;
            prof_mark(ctx, 1, so);
            k_observe<<<(n + wpb - 1) / wpb, OBS_THREADS, wpb * sizeof(ObserveSmem), so>>>(n, pool, st, tb, ctx->par, out);
            prof_mark(ctx, 1, so);
            ctx->launches++;
        }
        if (image && !image_last) {
            prof_mark(ctx, 6, so);
            render::Camera *cams = ctx->d_cams + lo;
            uint8_t *screen = ctx->d_screen + (size_t)lo * render::SCREEN_BYTES;
            int *repaint = ctx->d_repaint + lo, *repaint_n = ctx->d_repaint_n + lo;   // list and counter of this env range
            CK(cudaMemsetAsync(repaint_n, 0, sizeof(int), so));
            k_render_camera<<<(n + 127) / 128, 128, 0, so>>>(n, pool, st, ctx->par, cams, ctx->d_episode, ctx->d_screen_key + lo, repaint, repaint_n);
            k_render_static<<<std::min(n, ctx->sm_count * 4), render::THREADS, sizeof(render::Smem), so>>>(pool, st, ctx->d_episode, cams, ctx->par, screen, ctx->d_screen_key + lo, repaint, repaint_n);
            k_render<<<n, render::THREADS, sizeof(render::SmemDyn) + ctx->render_pad, so>>>(n, st, cams, ctx->par, ctx->palette, screen, out.img, ctx->d_spanrec + (size_t)lo * render::SPANREC, ctx->render_traj_len, ctx->render_force_lattice);
            ctx->launches++;
            prof_mark(ctx, 6, so);
            ctx->launches++;
        }
        if (early_out) { int rc = copy_fields(ctx, early_out, BY_OBSERVE, so, lo, cnt); if (rc) return rc; }
        if (fork) CK(cudaEventRecord(lane.ev_observed, so));
        if (!ctx->in_host_step && lo == 0 && cnt == ctx->n) {  // hope_wait_observed: the observation of the whole batch is complete here
            CK(cudaEventRecord(ctx->ev_obs_done, so));
            ctx->obs_done_valid = true;
        }
    }
    if (stages & HOPE_STAGE_RS) {
        const size_t wo = (size_t)lo * MAXW;
        RsScratch rs{ctx->d_words + wo, ctx->d_ntry + lo, ctx->d_ncand + lo, ctx->d_item_base + lo, ctx->d_items + wo,
                     ctx->d_item_bad + wo, ctx->d_slots + wo, ctx->d_n_items + chunk_id};
        prof_mark(ctx, 2, s);
        CK(cudaMemsetAsync(rs.n_items, 0, sizeof(int), s));
        k_rs_enumerate<<<(n + ENUM_THREADS - 1) / ENUM_THREADS, ENUM_THREADS, 0, s>>>(n, pool, st, tb, rs, out);
        prof_mark(ctx, 2, s);
        // persistent grids: the item count only exists on the device, so both kernels stride over it
        prof_mark(ctx, 3, s);
        k_rs_walk<<<ctx->walk_blocks, 128, 0, s>>>(rs, tb, ctx->par);
        prof_mark(ctx, 3, s);
        prof_mark(ctx, 4, s);
        // (host API: one resident block per SM fewer, so that the k_observe ranges and k_pack_lidar, which the copies wait for,
        // find registers next to the persistent grid; measured 2.11 -> 2.03 ms per host step)
        k_rs_check<<<ctx->in_host_step ? ctx->host_check_blocks : ctx->check_blocks, CHK_WARPS * 32, 0, s>>>(pool, st, tb, rs, ctx->par);
        prof_mark(ctx, 4, s);
        prof_mark(ctx, 5, s);
        k_rs_select<<<(n + 127) / 128, 128, 0, s>>>(n, tb, rs, out);
        prof_mark(ctx, 5, s);
        ctx->launches += 4;
    }
    if (image_last) {
        prof_mark(ctx, 6, s);
        render::Camera *cams = ctx->d_cams + lo;
        uint8_t *screen = ctx->d_screen + (size_t)lo * render::SCREEN_BYTES;
        int *repaint = ctx->d_repaint + lo, *repaint_n = ctx->d_repaint_n + lo;   // list and counter of this env range
        CK(cudaMemsetAsync(repaint_n, 0, sizeof(int), s));
        k_render_camera<<<(n + 127) / 128, 128, 0, s>>>(n, pool, st, ctx->par, cams, ctx->d_episode, ctx->d_screen_key + lo, repaint, repaint_n);
        k_render_static<<<std::min(n, ctx->sm_count * 4), render::THREADS, sizeof(render::Smem), s>>>(pool, st, ctx->d_episode, cams, ctx->par, screen, ctx->d_screen_key + lo, repaint, repaint_n);
        k_render<<<n, render::THREADS, sizeof(render::SmemDyn) + ctx->render_pad, s>>>(n, st, cams, ctx->par, ctx->palette, screen, out.img, ctx->d_spanrec + (size_t)lo * render::SPANREC, ctx->render_traj_len, ctx->render_force_lattice);
        prof_mark(ctx, 6, s);
        ctx->launches += 2;
    }
    if (fork) CK(cudaStreamWaitEvent(s, lane.ev_observed, 0));
    CK(cudaGetLastError());
    return HOPE_OK;
}

int launch_step(hope_ctx *ctx, const double *d_action, const hope_out &out, unsigned stages, int reset_all, cudaStream_t s) {
    return launch_range(ctx, d_action, out, stages, reset_all, s, 0, 0, 0, ctx->n);
}

// device API epilogue: remember where the caller's stream is, for a host-API call that may follow
int mark_device_call(hope_ctx *ctx, cudaStream_t s) {
    CK(cudaEventRecord(ctx->ev_dev, s));
    ctx->dev_pending = true;
    return HOPE_OK;
}
// host API prologue: order the context's own streams behind the last device-API call
int join_device_calls(hope_ctx *ctx, cudaStream_t s) {
    if (ctx->dev_pending) CK(cudaStreamWaitEvent(s, ctx->ev_dev, 0));
    return HOPE_OK;
}

// the static-screen cache of the image stage (62.5 KB per env), with the first image; never inside a stream capture
int ensure_screen(hope_ctx *ctx) {
    if (ctx->d_screen) return HOPE_OK;
    CK(cudaMalloc(&ctx->d_screen, (size_t)ctx->n * render::SCREEN_BYTES));
    CK(cudaMalloc(&ctx->d_screen_key, sizeof(uint2) * (size_t)ctx->n));
    CK(cudaMalloc(&ctx->d_spanrec, (size_t)ctx->n * render::SPANREC));
    CK(cudaMemset(ctx->d_spanrec, 0xff, (size_t)ctx->n * render::SPANREC));  // NaN poses: no record matches
    CK(cudaMalloc(&ctx->d_repaint, sizeof(int) * (size_t)ctx->n));
    CK(cudaMalloc(&ctx->d_repaint_n, sizeof(int) * (size_t)ctx->n));
    CK(cudaMemset(ctx->d_screen_key, 0xff, sizeof(uint2) * (size_t)ctx->n));
    CK(cudaDeviceSynchronize());  // the steps run on non-blocking streams, which do not order after the memset
    return HOPE_OK;
}
// the scene pool changed under the cached screens
int invalidate_screens(hope_ctx *ctx) {
    if (!ctx->d_screen_key) return HOPE_OK;
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(ctx->d_screen_key, 0xff, sizeof(uint2) * (size_t)ctx->n));
    CK(cudaDeviceSynchronize());
    return HOPE_OK;
}

// staging buffers of the host API; the image (12 KB per env) only when a caller asks for it
int ensure_stage(hope_ctx *ctx, bool want_img) {
    if (want_img) { const int rcs = ensure_screen(ctx); if (rcs) return rcs; }
    if (want_img && !ctx->d_stage_img) {
        CK(cudaMalloc(&ctx->d_stage_img, (size_t)ctx->n * HOPE_IMG_C * HOPE_IMG_HW * HOPE_IMG_HW));
        ctx->stage_out.img = static_cast<uint8_t *>(ctx->d_stage_img);
    }
    if (ctx->d_stage) return HOPE_OK;
    const size_t img_offset = offsetof(hope_out, img);
    size_t total = 0;
    for (int k = 0; k < kNumOutFields; ++k)
        if (kOutFields[k].offset != img_offset) total += ((kOutFields[k].elem * kOutFields[k].per_env * ctx->n + 255) / 256) * 256;
    CK(cudaMalloc(&ctx->d_stage, total));
    CK(cudaMemset(ctx->d_stage, 0, total));
    ctx->stage_bytes = total;
    size_t off = 0;
    for (int k = 0; k < kNumOutFields; ++k) {
        if (kOutFields[k].offset == img_offset) continue;
        field_ptr(ctx->stage_out, kOutFields[k]) = static_cast<char *>(ctx->d_stage) + off;
        off += ((kOutFields[k].elem * kOutFields[k].per_env * ctx->n + 255) / 256) * 256;
    }
    CK(cudaMalloc(&ctx->d_action, sizeof(double) * 2 * ctx->n));
    CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    return HOPE_OK;
}

}  // namespace

extern "C" {

int hope_version(void) { return 120; }  // 120: HOPE_STAGE_RAW_ACTION, NULL action = step without motion, env_collide, host/device API ordering
int hope_max_obs(void) { return HOPE_MAX_OBS; }

int hope_default_params(hope_params *p) {
    if (!p) return HOPE_ERR_INVALID;
    memset(p, 0, sizeof(*p));
    p->wheel_base = 2.8;
    const double front_hang = 0.96, rear_hang = 0.93, width = 1.94;
    p->box_x[0] = -rear_hang; p->box_x[1] = front_hang + p->wheel_base; p->box_x[2] = front_hang + p->wheel_base; p->box_x[3] = -rear_hang;
    p->box_y[0] = -width / 2; p->box_y[1] = -width / 2; p->box_y[2] = width / 2; p->box_y[3] = width / 2;
    p->valid_speed[0] = -2.5; p->valid_speed[1] = 2.5;
    p->valid_steer[0] = -0.75; p->valid_steer[1] = 0.75;
    p->num_step = 10; p->step_length = 0.05; p->mini_iter = 20;
    p->lidar_range = 10.0; p->tolerant_time = 200; p->rs_max_dist = 10.0; p->rs_step = 0.1;
    p->reward_weight[0] = 1; p->reward_weight[1] = 0; p->reward_weight[2] = 5; p->reward_weight[3] = 0; p->reward_weight[4] = 10;
    p->reward_ratio = 0.1; p->env_collide = 0; p->auto_reset = 1;
    p->regen_on_reset = 0; p->regen_level = -1; p->regen_seed = 20240529;
    return HOPE_OK;
}

const char *hope_strerror(int status) {
    switch (status) {
    case HOPE_OK: return "ok";
    case HOPE_ERR_INVALID: return "invalid argument";
    case HOPE_ERR_CUDA: return "CUDA runtime error (see hope_last_cuda_error)";
    case HOPE_ERR_NO_TABLES: return "tables not uploaded (hope_upload_tables)";
    case HOPE_ERR_NO_SCENES: return "scene pool empty or envs not reset (hope_set_scene_pool / hope_reset)";
    case HOPE_ERR_CAPACITY: return "capacity exceeded: a scene's HOPE_MAX_OBS / HOPE_MAX_VERTS, or the image stage's span table (hope_last_cuda_error)";
    default: return "unknown status";
    }
}

const char *hope_last_cuda_error(const hope_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
int hope_n_envs(const hope_ctx *ctx) { return ctx ? ctx->n : HOPE_ERR_INVALID; }

int hope_create(hope_ctx **out, int device, int n_envs, int pool_size, const hope_params *p) {
    if (!out || n_envs <= 0 || pool_size <= 0) return HOPE_ERR_INVALID;
    hope_ctx *ctx = new (std::nothrow) hope_ctx();
    if (!ctx) return HOPE_ERR_INVALID;
    *out = ctx;
    ctx->device = device; ctx->n = n_envs; ctx->pool = pool_size;
    if (p) ctx->par = *p; else hope_default_params(&ctx->par);
    if (ctx->par.regen_on_reset && pool_size < n_envs) { ctx->last_error = "regen_on_reset needs pool_size >= n_envs (env i owns slot i)"; return HOPE_ERR_INVALID; }
    memset(&ctx->stage_out, 0, sizeof(ctx->stage_out));
    ctx->maxc = tan(ctx->par.valid_steer[1]) / ctx->par.wheel_base;  // car_parking_base.py:422
    CK(cudaSetDevice(device));
    const size_t P = pool_size, N = n_envs;
    CK(cudaMalloc(&ctx->d_obs, sizeof(double) * P * MAXE * 2));
    CK(cudaMalloc(&ctx->d_aabb, sizeof(double) * P * MAXO * 4));
    CK(cudaMalloc(&ctx->d_meta, sizeof(double) * P * META));
    CK(cudaMalloc(&ctx->d_nv, P * MAXO));
    CK(cudaMalloc(&ctx->d_nobs, sizeof(int) * P));
    CK(cudaMemset(ctx->d_nv, 0, P * MAXO));
    CK(cudaMemset(ctx->d_nobs, 0, sizeof(int) * P));
    const size_t tab = 512 + 2 * (size_t)NUP * NACT * NITER + NUP + NRAY;
    CK(cudaMalloc(&ctx->d_tab, sizeof(double) * tab));
    CK(cudaMalloc(&ctx->d_pose, sizeof(double) * 3 * N));
    CK(cudaMalloc(&ctx->d_accum, sizeof(double) * N));
    CK(cudaMalloc(&ctx->d_cs, sizeof(double) * 2 * N));
    CK(cudaMalloc(&ctx->d_t, sizeof(int) * N));
    CK(cudaMalloc(&ctx->d_scene, sizeof(int) * N));
    CK(cudaMalloc(&ctx->d_pending, N));
    CK(cudaMalloc(&ctx->d_gate, N));
    CK(cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * 8));
    CK(cudaMemset(ctx->d_counters, 0, sizeof(unsigned long long) * 8));
    CK(cudaMemset(ctx->d_pending, 0, N));
    CK(cudaMemset(ctx->d_gate, 0, N));
    CK(cudaMalloc(&ctx->d_words, sizeof(RsWord) * N * MAXW));
    CK(cudaMalloc(&ctx->d_ntry, N));
    CK(cudaMalloc(&ctx->d_ncand, N));
    CK(cudaMalloc(&ctx->d_item_base, sizeof(int) * N));
    CK(cudaMalloc(&ctx->d_items, sizeof(int) * N * MAXW));
    CK(cudaMalloc(&ctx->d_item_bad, N * MAXW));
    CK(cudaMalloc(&ctx->d_slots, sizeof(WordSlot) * N * MAXW));
    CK(cudaMalloc(&ctx->d_n_items, sizeof(int) * 64));
    CK(cudaMalloc(&ctx->d_regen_slots, sizeof(int) * N));
    CK(cudaMalloc(&ctx->d_regen_count, sizeof(int)));
    CK(cudaMalloc(&ctx->d_gen_status, sizeof(int)));
    CK(cudaMemset(ctx->d_gen_status, 0, sizeof(int)));
    CK(cudaMalloc(&ctx->d_episode, sizeof(unsigned) * P));
    CK(cudaMemset(ctx->d_episode, 0, sizeof(unsigned) * P));
    CK(cudaMalloc(&ctx->d_traj, sizeof(double) * 80 * N));
    CK(cudaMalloc(&ctx->d_cams, sizeof(render::Camera) * N));
    CK(cudaMalloc(&ctx->d_traj_n, sizeof(int) * N));
    CK(cudaMemset(ctx->d_traj, 0, sizeof(double) * 80 * N));
    CK(cudaMemset(ctx->d_traj_n, 0, sizeof(int) * N));
    CK(cudaMalloc(&ctx->d_plan_rem, sizeof(double) * 5 * N));
    CK(cudaMalloc(&ctx->d_plan_u8, 9 * N));
    CK(cudaMemset(ctx->d_plan_u8, 0, 9 * N));
    CK(cudaMemset(ctx->d_n_items, 0, sizeof(int) * 64));
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->sm_count = v; }
    {   // persistent grids = exactly the number of co-resident blocks (multiples of the SM count)
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_rs_walk, 128, 0));
        ctx->walk_blocks = ctx->sm_count * (nb > 0 ? nb : 1);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_rs_check, CHK_WARPS * 32, 0));
        ctx->check_blocks = ctx->sm_count * (nb > 0 ? nb : 1);
        // tuning experiments: fewer resident Reeds-Shepp blocks per SM leave registers for k_observe's blocks
        ctx->host_check_blocks = ctx->sm_count * (nb > 1 ? nb - 1 : 1);
        if (const char *e = getenv("HOPE_B200_CHECK_BPS")) { int v = atoi(e); if (v >= 1 && v <= 8) ctx->check_blocks = ctx->host_check_blocks = ctx->sm_count * v; }
        if (const char *e = getenv("HOPE_B200_HOST_CHECK_BPS")) { int v = atoi(e); if (v >= 1 && v <= 8) ctx->host_check_blocks = ctx->sm_count * v; }
        if (const char *e = getenv("HOPE_B200_WALK_BPS")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->walk_blocks = ctx->sm_count * v; }
    }
    for (auto &ln : ctx->lanes) {
        CK(cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking));
        {   // the auxiliary stream carries k_observe, whose outputs are 90 % of the bytes the host API copies back:
            // give it the higher priority so its blocks are scheduled ahead of the Reeds-Shepp kernels and the
            // device-to-host copies start as early as possible (HOPE_B200_AUX_PRIORITY=0 turns this off)
            int lo_p = 0, hi_p = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            const char *e = getenv("HOPE_B200_AUX_PRIORITY");
            const bool prio = e ? atoi(e) != 0 : true;
            CK(cudaStreamCreateWithPriority(&ln.aux, cudaStreamNonBlocking, prio ? hi_p : lo_p));
        }
        CK(cudaEventCreateWithFlags(&ln.ev_advanced, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ln.ev_observed, cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_dev, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_obs_done, cudaEventDisableTiming));
    for (int li = 0; li < hope_ctx::MAX_LANES; ++li) CK(cudaEventCreateWithFlags(&ctx->ev_join[li], cudaEventDisableTiming));
    for (auto &e : ctx->ev_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ctx->ev_adv) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char *e = getenv("HOPE_B200_HOST_SPLIT_ADVANCE")) ctx->host_split_advance = atoi(e) != 0;
    memset(&ctx->hg_out, 0, sizeof(ctx->hg_out));
    if (const char *e = getenv("HOPE_B200_HOST_GRAPH")) ctx->host_graph_enabled = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_ZERO_COPY")) ctx->zero_copy_enabled = atoi(e) != 0;
    memset(&ctx->step_out, 0, sizeof(ctx->step_out));
    if (const char *e = getenv("HOPE_B200_DEVICE_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->device_chunks = v; }
    if (const char *e = getenv("HOPE_B200_RENDER_AFTER_RS")) ctx->render_after_rs = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_HOST_MASK_EXPAND")) ctx->host_mask_expand = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_HOST_LIDAR_PACK")) ctx->host_lidar_pack = atoi(e) != 0;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) { int v = atoi(e); if (v >= 1) ctx->ranks_per_box = v; }  // torchrun: ranks sharing this host
    ctx->pack_frac = ctx->ranks_per_box == 1 ? 1.0 : (ctx->ranks_per_box == 2 ? 0.75 : 0.0);  // measured on B200 boxes with 32 CPUs (profiles/r02_e2e_ranks{2,4,8}.jsonl): 2 ranks 1.0 -> 2.24 ms per step, 0.75 -> 2.08, 0.5 -> 2.27, 0 -> 2.63; 4 ranks 1.0 -> 4.49, 0.5 -> 3.51, 0.25 -> 3.43, 0 -> 3.16; 8 ranks 1.0 -> 9.8, 0.25 -> 7.7, 0 -> 8.2 (7.7 - 8.2 is the spread between runs)
    if (const char *e = getenv("HOPE_B200_HOST_PACK_FRAC")) { const double v = atof(e); if (v >= 0.0 && v <= 1.0) ctx->pack_frac = v; }
    if (ctx->pack_frac <= 0.0) ctx->host_lidar_pack = false;   // nothing would travel packed: skip k_pack_lidar and its arrays altogether
    if (const char *e = getenv("HOPE_B200_WIRE_PORTABLE")) ctx->wire_force_portable = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_HOST_NOEXPAND")) ctx->host_noexpand = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_HOST_TRACE")) ctx->host_trace = atoi(e);
    if (const char *e = getenv("HOPE_B200_HOST_SPLIT")) ctx->host_split = e;
    if (const char *e = getenv("HOPE_B200_HOST_DEBUG")) ctx->host_debug = atoi(e);
    if (const char *e = getenv("HOPE_B200_RENDER_LATTICE")) ctx->render_force_lattice = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_RENDER_PAD_KB")) { int v = atoi(e); if (v > 0 && v <= 150) ctx->render_pad = (size_t)v * 1024; }
    if (const char *e = getenv("HOPE_B200_HOST_RS_AFTER")) ctx->host_rs_after_observe = atoi(e) != 0;
    if (const char *e = getenv("HOPE_B200_HOST_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= 64) ctx->host_chunks = v; }
    CK(cudaFuncSetAttribute(k_observe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((OBS_THREADS / 32) * sizeof(ObserveSmem))));
    CK(cudaFuncSetAttribute(k_render_static, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(render::Smem)));
    CK(cudaFuncSetAttribute(k_render, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(render::SmemDyn) + ctx->render_pad)));
    {   // a vehicle-sized box covers at most diagonal * K + 2 screen rows after truncation
        double diag = 0.0;
        for (int a = 0; a < 4; ++a)
            for (int b = a + 1; b < 4; ++b) diag = fmax(diag, hypot(ctx->par.box_x[a] - ctx->par.box_x[b], ctx->par.box_y[a] - ctx->par.box_y[b]));
        ctx->image_ok = diag * render::KSCALE + 2.0 <= render::DROWS;
    }
    {   // configs.py:26-30, 80-88; TRAJ_COLORS = np.linspace(LOW, HIGH, 20, endpoint=True, dtype=np.uint8)
        uint8_t rgb[HOPE_N_COLOR][3] = {{255, 255, 255}, {150, 150, 150}, {100, 149, 237}, {69, 139, 0}, {30, 144, 255}};
        const double low[3] = {10, 10, 10}, high[3] = {10, 10, 200};
        for (int k = 0; k < 20; ++k)
            for (int c = 0; c < 3; ++c) {
                const double step = (high[c] - low[c]) / 19;
                rgb[5 + k][c] = (uint8_t)(k == 19 ? high[c] : low[c] + k * step);
            }
        hope_set_palette(ctx, &rgb[0][0]);
    }
    return HOPE_OK;
}

int hope_destroy(hope_ctx *ctx) {
    if (!ctx) return HOPE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    void *ptrs[] = {ctx->d_obs, ctx->d_aabb, ctx->d_meta, ctx->d_nv, ctx->d_nobs, ctx->d_tab, ctx->d_pose, ctx->d_accum, ctx->d_t,
                    ctx->d_scene, ctx->d_pending, ctx->d_gate, ctx->d_counters, ctx->d_words, ctx->d_ntry, ctx->d_ncand, ctx->d_cs, ctx->d_item_base, ctx->d_items, ctx->d_item_bad, ctx->d_slots, ctx->d_n_items, ctx->d_plan_rem, ctx->d_plan_u8, ctx->d_regen_slots, ctx->d_regen_count, ctx->d_gen_status, ctx->d_episode,
                    ctx->d_action, ctx->d_stage, ctx->d_stage_img, ctx->d_traj, ctx->d_traj_n, ctx->d_cams, ctx->d_screen, ctx->d_screen_key, ctx->d_repaint, ctx->d_repaint_n, ctx->d_spanrec};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->host_graph) cudaGraphExecDestroy(ctx->host_graph);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_dev) cudaEventDestroy(ctx->ev_dev);
    if (ctx->ev_obs_done) cudaEventDestroy(ctx->ev_obs_done);
    for (auto e : ctx->ev_join) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_adv) if (e) cudaEventDestroy(e);
    if (ctx->h_mask_steps) cudaFreeHost(ctx->h_mask_steps);
    for (void *p : {(void *)ctx->d_lbits, (void *)ctx->d_loff, (void *)ctx->d_lpacked, (void *)ctx->d_lcount}) if (p) cudaFree(p);
    for (void *p : {(void *)ctx->h_lbits, (void *)ctx->h_loff, (void *)ctx->h_lpacked, (void *)ctx->h_lcount}) if (p) cudaFreeHost(p);
    if (ctx->s_pack) cudaStreamDestroy(ctx->s_pack);
    for (auto e : ctx->ev_pack) if (e) cudaEventDestroy(e);
    if (ctx->h_seq) cudaFreeHost(const_cast<unsigned long long *>(ctx->h_seq));
    if (ctx->d_seq) cudaFree(ctx->d_seq);
    delete ctx->host_pool;
    for (auto &ln : ctx->lanes) {
        if (ln.main) cudaStreamDestroy(ln.main);
        if (ln.aux) cudaStreamDestroy(ln.aux);
        if (ln.ev_advanced) cudaEventDestroy(ln.ev_advanced);
        if (ln.ev_observed) cudaEventDestroy(ln.ev_observed);
    }
    delete ctx;
    return HOPE_OK;
}

int hope_set_palette(hope_ctx *ctx, const uint8_t *h_rgb) {
    if (!ctx || !h_rgb) return HOPE_ERR_INVALID;
    const uint8_t *bg = h_rgb;
    for (int k = 0; k < HOPE_N_COLOR; ++k) {
        const uint8_t *c = h_rgb + 3 * k;
        const bool white = c[0] == bg[0] && c[1] == bg[1] && c[2] == bg[2];  // change_bg_color: BG_COLOR pixels become black
        ctx->palette.rg[k] = white ? 0u : ((uint32_t)c[0] | ((uint32_t)c[1] << 16));
        ctx->palette.b[k] = white ? 0u : (uint32_t)c[2];
    }
    // k_render takes the palette by value: a captured host step would keep painting with the old colours
    if (ctx->host_graph) { cudaGraphExecDestroy(ctx->host_graph); ctx->host_graph = nullptr; ctx->hg_action = nullptr; }
    return HOPE_OK;
}

int hope_set_render_traj(hope_ctx *ctx, int traj_render_len) {
    if (!ctx || traj_render_len < 0 || traj_render_len > render::TRAJ) return HOPE_ERR_INVALID;
    ctx->render_traj_len = traj_render_len;
    // k_render takes it by value: a captured host step would keep the old length
    if (ctx->host_graph) { cudaGraphExecDestroy(ctx->host_graph); ctx->host_graph = nullptr; ctx->hg_action = nullptr; }
    return HOPE_OK;
}

int hope_upload_tables(hope_ctx *ctx, const double *ray_a, const double *ray_b, const double *lidar_base, const double *mask_base,
                       const double *dist_star, const double *w_lo, const double *w_hi) {
    if (!ctx || !ray_a || !ray_b || !lidar_base || !mask_base || !dist_star || !w_lo || !w_hi) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> head(512, 0.0);
    memcpy(&head[0], ray_a, 120 * 8); memcpy(&head[120], ray_b, 120 * 8); memcpy(&head[240], lidar_base, 120 * 8);
    memcpy(&head[360], mask_base, 120 * 8); memcpy(&head[480], w_lo, 80); memcpy(&head[496], w_hi, 80);
    CK(cudaMemcpy(ctx->d_tab, head.data(), 512 * 8, cudaMemcpyHostToDevice));
    for (int j = 0; j < NRAY; ++j) ctx->h_nohit[j] = ctx->par.lidar_range - lidar_base[j];
    double *ds = ctx->d_tab + 512;
    CK(cudaMemcpy(ds, dist_star, sizeof(double) * NUP * NACT * NITER, cudaMemcpyHostToDevice));
    double *pmaxk = ds + (size_t)NUP * NACT * NITER, *pmax = pmaxk + (size_t)NUP * NACT * NITER;
    k_table_reduce<<<NUP, 64>>>(ds, pmaxk, pmax);
    k_table_group<<<1, 128>>>(pmax, pmax + NUP);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    ctx->have_tables = true;
    return HOPE_OK;
}

int hope_set_scene_pool(hope_ctx *ctx, int first, int n, const double *start, const double *dest, const double *bounds,
                        const double *obs_xy, const int32_t *nverts) {
    if (!ctx || first < 0 || n <= 0 || first + n > ctx->pool || !start || !dest || !bounds || !obs_xy || !nverts) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> meta((size_t)n * META), aabb((size_t)n * MAXO * 4, 0.0), obs((size_t)n * MAXE * 2, 0.0);
    std::vector<uint8_t> nv((size_t)n * MAXO, 0);
    std::vector<int> nobs(n, 0);
    const hope_params &par = ctx->par;
    for (int i = 0; i < n; ++i) {
        double *m = &meta[(size_t)i * META];
        for (int k = 0; k < 3; ++k) { m[M_START + k] = start[3 * i + k]; m[M_DEST + k] = dest[3 * i + k]; }
        for (int k = 0; k < 4; ++k) m[M_BOUNDS + k] = bounds[4 * i + k];
        // dest box (State.create_box, vehicle.py:32-36) and its shoelace area, on the host libm
        double c = cos(dest[3 * i + 2]), s = sin(dest[3 * i + 2]), ms = -s;
        double bx[4], by[4];
        for (int k = 0; k < 4; ++k) {
            bx[k] = c * par.box_x[k] + ms * par.box_y[k] + dest[3 * i];
            by[k] = s * par.box_x[k] + c * par.box_y[k] + dest[3 * i + 1];
            m[M_DBX + k] = bx[k]; m[M_DBY + k] = by[k];
        }
        double sa = 0.0;
        for (int k = 0; k < 4; ++k) { int j = (k + 1) & 3; sa += bx[k] * by[j] - bx[j] * by[k]; }
        m[M_DAREA] = fabs(sa) * 0.5;
        m[M_DNORM] = fmax(hypot(dest[3 * i] - start[3 * i], dest[3 * i + 1] - start[3 * i + 1]), 10.0);  // car_parking_base.py:211
        m[M_DAABB] = fmin(fmin(bx[0], bx[1]), fmin(bx[2], bx[3])); m[M_DAABB + 1] = fmax(fmax(bx[0], bx[1]), fmax(bx[2], bx[3]));
        m[M_DAABB + 2] = fmin(fmin(by[0], by[1]), fmin(by[2], by[3])); m[M_DAABB + 3] = fmax(fmax(by[0], by[1]), fmax(by[2], by[3]));
        // obstacles: compact non-empty rings to the front
        int no = 0;
        for (int k = 0; k < MAXO; ++k) {
            int v = nverts[(size_t)i * MAXO + k];
            if (v == 0) continue;
            if (v < 3 || v > MAXV) return HOPE_ERR_CAPACITY;
            const double *src = obs_xy + ((size_t)i * MAXO + k) * MAXV * 2;
            double *dst = &obs[((size_t)i * MAXO + no) * MAXV * 2];
            double xmn = src[0], xmx = src[0], ymn = src[1], ymx = src[1];
            for (int j = 0; j < v; ++j) {
                dst[2 * j] = src[2 * j]; dst[2 * j + 1] = src[2 * j + 1];
                xmn = fmin(xmn, src[2 * j]); xmx = fmax(xmx, src[2 * j]); ymn = fmin(ymn, src[2 * j + 1]); ymx = fmax(ymx, src[2 * j + 1]);
            }
            double *bb = &aabb[((size_t)i * MAXO + no) * 4];
            bb[0] = xmn; bb[1] = xmx; bb[2] = ymn; bb[3] = ymx;
            nv[(size_t)i * MAXO + no] = (uint8_t)v;
            ++no;
        }
        nobs[i] = no;
    }
    CK(cudaMemcpy(ctx->d_meta + (size_t)first * META, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_aabb + (size_t)first * MAXO * 4, aabb.data(), aabb.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_obs + (size_t)first * MAXE * 2, obs.data(), obs.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_nv + (size_t)first * MAXO, nv.data(), nv.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_nobs + first, nobs.data(), nobs.size() * sizeof(int), cudaMemcpyHostToDevice));
    ctx->have_scenes = true;
    return invalidate_screens(ctx);
}

int hope_generate_scene_pool_device(hope_ctx *ctx, int first, int n, int level, uint64_t seed, void *stream) {
    if (!ctx || first < 0 || n <= 0 || first + n > ctx->pool || level < -1 || level > 2) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (hope_scene::launch_generate(n, first, level, seed, nullptr, nullptr, ctx->par, ctx->d_obs, ctx->d_nv, ctx->d_aabb, ctx->d_meta, ctx->d_nobs,
                                    ctx->d_gen_status, stream) != HOPE_OK) return fail(ctx, cudaGetLastError(), "k_generate_scenes");
    ctx->launches++;
    ctx->have_scenes = true;
    return invalidate_screens(ctx);
}

int hope_get_scene_pool(hope_ctx *ctx, int first, int n, double *h_start, double *h_dest, double *h_bounds, double *h_obs_xy, int32_t *h_nverts) {
    if (!ctx || first < 0 || n <= 0 || first + n > ctx->pool) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    int gs = 0;
    CK(cudaMemcpy(&gs, ctx->d_gen_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (gs != 0) return gs;  // a device generator thread gave up (HOPE_ERR_INVALID) or overflowed HOPE_MAX_OBS
    std::vector<double> meta((size_t)n * META);
    std::vector<uint8_t> nv((size_t)n * MAXO);
    CK(cudaMemcpy(meta.data(), ctx->d_meta + (size_t)first * META, meta.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nv.data(), ctx->d_nv + (size_t)first * MAXO, nv.size(), cudaMemcpyDeviceToHost));
    if (h_obs_xy) CK(cudaMemcpy(h_obs_xy, ctx->d_obs + (size_t)first * MAXE * 2, sizeof(double) * n * MAXE * 2, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const double *m = &meta[(size_t)i * META];
        for (int k = 0; k < 3; ++k) { if (h_start) h_start[3 * i + k] = m[M_START + k]; if (h_dest) h_dest[3 * i + k] = m[M_DEST + k]; }
        for (int k = 0; k < 4; ++k) if (h_bounds) h_bounds[4 * i + k] = m[M_BOUNDS + k];
        for (int k = 0; k < MAXO; ++k) if (h_nverts) h_nverts[(size_t)i * MAXO + k] = nv[(size_t)i * MAXO + k];
    }
    return HOPE_OK;
}

int hope_reset(hope_ctx *ctx, const int32_t *h_scene_ids, const hope_out *d_out, void *stream) {
    if (!ctx || !d_out) return HOPE_ERR_INVALID;
    if (!ctx->have_tables) return HOPE_ERR_NO_TABLES;
    if (!ctx->have_scenes) return HOPE_ERR_NO_SCENES;
    CK(cudaSetDevice(ctx->device));
    std::vector<int> ids(ctx->n);
    for (int i = 0; i < ctx->n; ++i) {
        ids[i] = h_scene_ids ? h_scene_ids[i] : i % ctx->pool;
        if (ids[i] < 0 || ids[i] >= ctx->pool) return HOPE_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(ctx->d_scene, ids.data(), sizeof(int) * ctx->n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // ids is a local buffer
    ctx->have_reset = true;
    if (d_out->img) { const int rcs = ensure_screen(ctx); if (rcs) return rcs; }
    // the reset step computes the observation; RS is gated off by t > 1 (car_parking_base.py:293)
    int rc = launch_step(ctx, nullptr, *d_out, HOPE_STAGE_ADVANCE | HOPE_STAGE_OBSERVE | HOPE_STAGE_RS | HOPE_STAGE_IMAGE, 1, s);
    if (rc) return rc;
    return mark_device_call(ctx, s);
}

int hope_step(hope_ctx *ctx, const double *d_action, const hope_out *d_out, unsigned stages, void *stream) {
    if (!ctx || !d_out) return HOPE_ERR_INVALID;  // d_action == NULL: CarParking.step(None), no motion (car_parking_base.py:255)
    if (!ctx->have_tables) return HOPE_ERR_NO_TABLES;
    if (!ctx->have_reset) return HOPE_ERR_NO_SCENES;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    stages |= HOPE_STAGE_ADVANCE;
    if ((stages & HOPE_STAGE_IMAGE) && d_out->img) { const int rcs = ensure_screen(ctx); if (rcs) return rcs; }
    ctx->obs_done_valid = false;
    const int n = ctx->n;
    int chunks = ctx->device_chunks;
    if (n < 8192 * chunks) chunks = n / 8192 > 0 ? n / 8192 : 1;
    if (chunks <= 1) {
        int rc1 = launch_step(ctx, d_action, *d_out, stages, 0, s);
        return rc1 ? rc1 : mark_device_call(ctx, s);
    }
    // fork the lanes from the caller's stream, one env range per lane (round robin), join back
    const int per = ((n + chunks - 1) / chunks + 127) / 128 * 128;
    CK(cudaEventRecord(ctx->ev_fork, s));
    for (int li = 0; li < hope_ctx::MAX_LANES; ++li) CK(cudaStreamWaitEvent(ctx->lanes[li].main, ctx->ev_fork, 0));
    for (int c = 0, lo = 0; lo < n; ++c, lo += per) {
        const int cnt = (lo + per <= n) ? per : n - lo;
        const int li = c % hope_ctx::MAX_LANES;
        int rc = launch_range(ctx, d_action, *d_out, stages, 0, ctx->lanes[li].main, li, c, lo, cnt);
        if (rc) return rc;
    }
    for (int li = 0; li < hope_ctx::MAX_LANES; ++li) {
        CK(cudaEventRecord(ctx->ev_join[li], ctx->lanes[li].main));
        CK(cudaStreamWaitEvent(s, ctx->ev_join[li], 0));
    }
    return mark_device_call(ctx, s);
}

int hope_step_kinematics_collision(hope_ctx *ctx, const double *d_action, double *d_pose, uint8_t *d_collided, uint8_t *d_substeps,
                                   void *stream) {
    if (!ctx || !d_action) return HOPE_ERR_INVALID;
    if (!ctx->have_reset) return HOPE_ERR_NO_SCENES;
    CK(cudaSetDevice(ctx->device));
    hope_out o;
    memset(&o, 0, sizeof(o));
    o.pose = d_pose; o.retreated = d_collided; o.substeps = d_substeps;
    int rc = launch_step(ctx, d_action, o, HOPE_STAGE_ADVANCE, 0, static_cast<cudaStream_t>(stream));
    return rc ? rc : mark_device_call(ctx, static_cast<cudaStream_t>(stream));
}

// enqueue one pipelined host step (see hope_step_host) on the context's lanes; lane 0's main stream is the origin:
// the other lanes fork from it and join back, which also makes the whole thing capturable as one CUDA graph
static void plan_zero_copy(hope_ctx *ctx, const hope_host_out *h_out) {
    ctx->step_out = ctx->stage_out;
    ctx->zero_copy_mask = 0;
    if (!ctx->zero_copy_enabled) return;
    for (int k = 0; k < kNumOutFields; ++k) {
        void *h = field_ptr_c(*h_out, kOutFields[k]);
        if (!h) continue;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        if (at.type == cudaMemoryTypeHost && at.devicePointer) {
            field_ptr(ctx->step_out, kOutFields[k]) = at.devicePointer;
            ctx->zero_copy_mask |= 1u << k;
        }
    }
}

// Decide which arrays of this host step travel in the narrow wire format (see hope_ctx::host_mask_expand, host_lidar_pack).
static int plan_wire(hope_ctx *ctx, const hope_host_out *h_out, unsigned stages) {
    const int lidar_field = 1, mask_field = 2;  // indices of `lidar`, `mask` in kOutFields
    ctx->expanding = ctx->host_mask_expand && h_out->mask && (stages & HOPE_STAGE_OBSERVE) && !((ctx->zero_copy_mask >> mask_field) & 1);
    ctx->packing = ctx->host_lidar_pack && h_out->lidar && (stages & HOPE_STAGE_OBSERVE) && !((ctx->zero_copy_mask >> lidar_field) & 1);
    if (!ctx->expanding && !ctx->packing) return HOPE_OK;
    if (!ctx->h_seq) {
        unsigned long long *hs = nullptr;
        CK(cudaMallocHost(&hs, sizeof(unsigned long long) * 2 * hope_ctx::MAX_CHUNK_EVENTS));  // two flags per env range
        memset(hs, 0, sizeof(unsigned long long) * 2 * hope_ctx::MAX_CHUNK_EVENTS);
        ctx->h_seq = hs;
        CK(cudaMalloc(&ctx->d_seq, sizeof(unsigned long long)));
        CK(cudaMemset(ctx->d_seq, 0, sizeof(unsigned long long)));
        CK(cudaDeviceSynchronize());  // the step runs on non-blocking streams, which do not order after the memset
        ctx->host_seq = 0;
    }
    if (ctx->expanding && !ctx->h_mask_steps) CK(cudaMallocHost(&ctx->h_mask_steps, (size_t)ctx->n * NACT));
    if (ctx->packing && !ctx->d_lpacked) {
        const size_t N = ctx->n;
        CK(cudaMalloc(&ctx->d_lbits, sizeof(uint32_t) * 4 * N));
        CK(cudaMalloc(&ctx->d_loff, sizeof(uint32_t) * N));
        CK(cudaMalloc(&ctx->d_lpacked, sizeof(double) * NRAY * N));
        CK(cudaMalloc(&ctx->d_lcount, sizeof(unsigned) * hope_ctx::MAX_CHUNK_EVENTS));
        CK(cudaMallocHost(&ctx->h_lbits, sizeof(uint32_t) * 4 * N));
        CK(cudaMallocHost(&ctx->h_loff, sizeof(uint32_t) * N));
        CK(cudaMallocHost(&ctx->h_lpacked, sizeof(double) * NRAY * N + 64));  // + what the 8-wide expansion may read behind the last value
        CK(cudaMallocHost(&ctx->h_lcount, sizeof(unsigned) * hope_ctx::MAX_CHUNK_EVENTS));
        CK(cudaStreamCreateWithFlags(&ctx->s_pack, cudaStreamNonBlocking));
        for (auto &e : ctx->ev_pack) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!ctx->host_pool) {
        ctx->host_pool = new (std::nothrow) HostPool();
        if (!ctx->host_pool) return HOPE_ERR_INVALID;
        // default: the CPUs this process may run on, shared by the ranks of the node (torchrun's LOCAL_WORLD_SIZE), one left
        // to the stepping thread; HOPE_B200_HOST_THREADS overrides
        int cpus = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
        int nt = cpus / ctx->ranks_per_box - 1;
        nt = nt < 1 ? 1 : (nt > 12 ? 12 : nt);
        if (const char *e = getenv("HOPE_B200_HOST_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) nt = v; }
        ctx->host_threads = nt;
        if (const char *e = getenv("HOPE_B200_HOST_SPIN_US")) ctx->host_pool->spin_us = atoi(e);
        ctx->host_pool->start(nt);
    }
    if (ctx->expanding) ctx->step_out.mask = nullptr;  // k_observe does not write the float64 mask at all
    return HOPE_OK;
}

// After the step has been launched: the stepping thread watches the per-range flags.  When range c's narrow arrays have
// landed it issues the copy of the range's kept lidar values (now that their count is known) and queues the mask expansion;
// when that copy has finished it queues the lidar expansion.  Then it waits for the workers and for the rest of the step.
static int finish_host_step(hope_ctx *ctx, const hope_host_out *h_out, cudaStream_t s0) {
    ctx->io_d2h = ctx->io_d2h_static;
    if (ctx->expanding || ctx->packing) {
        const unsigned long long expected = ++ctx->host_seq;  // k_bump_seq ran (or will run) once more on the device
        const int C = ctx->hm_chunks, portable = ctx->wire_force_portable;
        const int G = ctx->packing ? ctx->pk_total : 0;
        volatile unsigned long long *flag_a = ctx->h_seq, *flag_b = ctx->h_seq + hope_ctx::MAX_CHUNK_EVENTS;
        int next_a = ctx->packing ? 0 : C, next_b = 0, next_lidar = 0, issued = 0;
        for (unsigned spins = 0; next_a < C || next_b < C || next_lidar < G; ++spins) {
            if (next_a < C && flag_a[next_a] == expected) {  // the range's sub-range counts are here: copy the kept values
                const int c = next_a++;
                hmark(ctx, "counts_seen", c);
                for (int g = ctx->pk_first[c]; g < ctx->pk_first[c + 1]; ++g) {
                    const size_t glo = ctx->pk_lo[g], kept = ctx->h_lcount[g];
                    if (kept > (size_t)(ctx->pk_hi[g] - ctx->pk_lo[g]) * NRAY) { ctx->last_error = "hope_step_host: lidar pack count out of range"; return HOPE_ERR_CUDA; }
                    if (g == ctx->pk_first[c]) tmark(ctx, "kept_copy_begin", c, ctx->s_pack);
                    if (!ctx->pk_packed[g]) {  // this sub-range's float64 rows go straight into the caller's buffer
                        const size_t bytes = (size_t)(ctx->pk_hi[g] - ctx->pk_lo[g]) * NRAY * sizeof(double);
                        CK(cudaMemcpyAsync(h_out->lidar + glo * NRAY, ctx->stage_out.lidar + glo * NRAY, bytes, cudaMemcpyDeviceToHost, ctx->s_pack));
                        ctx->io_d2h += bytes;
                    } else if (kept) {
                        CK(cudaMemcpyAsync(ctx->h_lpacked + glo * NRAY, ctx->d_lpacked + glo * NRAY, kept * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_pack));
                        ctx->io_d2h += kept * sizeof(double);
                    }
                    CK(cudaEventRecord(ctx->ev_pack[g], ctx->s_pack));
                }
                tmark(ctx, "kept_copy_end", c, ctx->s_pack);
                hmark(ctx, "kept_copy_issued", c);
                continue;
            }
            if (next_b < C && next_b < (ctx->packing ? next_a : C) && flag_b[next_b] == expected) {  // step counts, beam flags and offsets are here
                const int c = next_b++;
                const size_t lo = ctx->ch_lo[c], hi = ctx->ch_lo[c + 1];
                hmark(ctx, "narrow_seen", c);
                issued = ctx->packing ? ctx->pk_first[c + 1] : 0;
                if (ctx->expanding && !ctx->host_noexpand) {
                    const uint8_t *steps = ctx->h_mask_steps;
                    double *mask = h_out->mask;
                    ctx->host_pool->submit([=](int part, int nparts) {
                        const size_t span = (hi - lo + nparts - 1) / nparts, a = lo + span * part, b = a + span < hi ? a + span : hi;
                        if (a < b) hope_wire::expand_mask(steps, mask, a, b, portable);
                    });
                }
                continue;
            }
            if (next_lidar < issued) {
                const cudaError_t q = cudaEventQuery(ctx->ev_pack[next_lidar]);
                if (q == cudaSuccess) {
                    const int g = next_lidar++;
                    hmark(ctx, "kept_copy_seen_done", g);
                    if (ctx->host_noexpand || !ctx->pk_packed[g]) continue;
                    const size_t lo = ctx->pk_lo[g], hi = ctx->pk_hi[g];
                    const uint32_t *bits = ctx->h_lbits, *off = ctx->h_loff;
                    const double *packed = ctx->h_lpacked + lo * NRAY, *nohit = ctx->h_nohit;
                    double *lidar = h_out->lidar;
                    ctx->host_pool->submit([=](int part, int nparts) {
                        const size_t span = (hi - lo + nparts - 1) / nparts, a = lo + span * part, b = a + span < hi ? a + span : hi;
                        if (a < b) hope_wire::expand_lidar(bits, off, packed, nohit, lidar, a, b, portable);
                    });
                    continue;
                }
                if (q != cudaErrorNotReady) return fail(ctx, q, "hope_step_host (lidar values)");
            }
            if ((spins & 0xfff) == 0xfff && (next_a < C || next_b < C)) {  // the step died or finished without delivering: do not spin forever
                const cudaError_t q = cudaStreamQuery(s0);
                if (q != cudaErrorNotReady) {
                    if (q != cudaSuccess) return fail(ctx, q, "hope_step_host");
                    if ((next_a < C && flag_a[next_a] != expected) || (next_b < C && flag_b[next_b] != expected)) {
                        ctx->last_error = "hope_step_host: narrow observation arrays did not arrive";
                        return HOPE_ERR_CUDA;
                    }
                }
            }
        }
        hmark(ctx, "all_jobs_queued", -1);
        ctx->host_pool->wait_all();
        hmark(ctx, "workers_done", -1);
    }
    CK(cudaStreamSynchronize(s0));
    hmark(ctx, "stream_synchronised", -1);
    return HOPE_OK;
}

// One host step, ordered so that the device-to-host copies start as early as possible and never wait for the
// Reeds-Shepp kernels: k_observe's outputs (lidar + mask) are 90 % of the bytes.
//   s0     : H2D(actions), k_advance over all envs, then the Reeds-Shepp kernels over all envs, then their small outputs
//   s_obs  : (high priority) k_observe [+ k_render] env range by env range, in order
//   s_copy : the observation buffers of a range, as soon as that range's k_observe is done
// s0 is the origin; the other two fork from it and join back, so the whole step is capturable as one CUDA graph.
// (Letting every range run advance -> observe -> RS on its own stream pair finishes all ranges at about the same
// time: the first copy then starts ~0.9 ms into the step whatever the number of ranges.)
static int enqueue_host_step(hope_ctx *ctx, const double *h_action, const hope_host_out *h_out, unsigned stages) {
    if (!h_out->img) stages &= ~(unsigned)HOPE_STAGE_IMAGE;  // nobody reads the staged image
    const unsigned side = stages & (HOPE_STAGE_OBSERVE | HOPE_STAGE_IMAGE);
    const int n = ctx->n;
    struct Scope { hope_ctx *c; ~Scope() { c->in_host_step = false; } } scope{ctx};
    ctx->in_host_step = true;
    cudaStream_t s0 = ctx->lanes[0].main, s_obs = ctx->lanes[0].aux, s_copy = ctx->lanes[1].main;
    const double *d_act = h_action ? ctx->d_action : nullptr;  // NULL: step without motion (CarParking.step(None))
    const unsigned adv = HOPE_STAGE_ADVANCE | (stages & HOPE_STAGE_RAW_ACTION);
    tmark(ctx, "start", -1, s0);
    if (h_action) CK(cudaMemcpyAsync(ctx->d_action, h_action, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s0));
    tmark(ctx, "actions_in", -1, s0);
    ctx->io_h2d = h_action ? sizeof(double) * 2 * (size_t)n : 0;
    ctx->io_d2h_static = 0;
    const bool flagging = ctx->expanding || ctx->packing;
    if (flagging) { k_bump_seq<<<1, 1, 0, s0>>>(ctx->d_seq); ctx->launches++; }
    int rc;
    // env ranges of the pipeline: host_chunks equal ones, or the percentages of HOPE_B200_HOST_SPLIT; boundaries on multiples of 128
    int C = 0;
    {
        int chunks = ctx->host_chunks;
        if (n < 4096 * chunks) chunks = n / 4096 > 0 ? n / 4096 : 1;
        std::vector<double> share;
        if (!ctx->host_split.empty() && n >= 8192) {
            const char *p = ctx->host_split.c_str();
            while (*p) { char *end = nullptr; const double v = strtod(p, &end); if (end == p) break; if (v > 0) share.push_back(v); p = *end ? end + 1 : end; }
            if ((int)share.size() > hope_ctx::MAX_CHUNK_EVENTS) share.clear();
        }
        if (share.empty()) share.assign(chunks, 1.0);
        double total = 0, acc = 0;
        for (double v : share) total += v;
        ctx->ch_lo[0] = 0;
        for (size_t k = 0; k < share.size(); ++k) {
            acc += share[k];
            int hi = k + 1 == share.size() ? n : (int)((double)n * acc / total + 127) / 128 * 128;
            if (hi > n) hi = n;
            if (hi > ctx->ch_lo[C]) ctx->ch_lo[++C] = hi;
        }
        if (ctx->ch_lo[C] != n) ctx->ch_lo[++C] = n;
    }
    const bool split_advance = side && ctx->host_split_advance;  // k_advance per range too: the first range's k_observe starts earlier
    if (!split_advance) {
        rc = launch_range(ctx, d_act, ctx->step_out, adv, 0, s0, 0, 0, 0, n);
        if (rc) return rc;
        tmark(ctx, "advance_end", -1, s0);
    }
    int last_chunk = 0;
    bool advance_copied = false;
    if (side) {
        if (!split_advance) {
            CK(cudaEventRecord(ctx->ev_fork, s0));
            CK(cudaStreamWaitEvent(s_obs, ctx->ev_fork, 0));
            CK(cudaStreamWaitEvent(s_copy, ctx->ev_fork, 0));
            rc = copy_fields(ctx, h_out, BY_ADVANCE, s_copy, 0, n);
            if (rc) return rc;
            advance_copied = true;
        }
        ctx->hm_chunks = C;
        {   // sub-ranges of the kept-value copies: at most 64 over all ranges
            int sub = 2048;  // 32 sub-ranges at 65 536 envs: the host work left behind the last copy is 1/32 of the expansion (B200, a box whose memory bounds the expansion: 8192 -> 2.09 ms per host step, 4096 -> 1.95, 2048 -> 1.88; equal on faster hosts)
            if (const char *e = getenv("HOPE_B200_PACK_SUB")) { int v = atoi(e); if (v >= 128) sub = v / 128 * 128; }
            for (;;) {
                int total = 0;
                for (int c = 0; c < C; ++c) total += (ctx->ch_lo[c + 1] - ctx->ch_lo[c] + sub - 1) / sub;
                if (total <= 64) break;
                sub *= 2;
            }
            ctx->pk_sub = sub;
            int g = 0;
            for (int c = 0; c < C; ++c) {
                const int lo = ctx->ch_lo[c], cnt = ctx->ch_lo[c + 1] - lo;
                ctx->pk_first[c] = g;
                for (int a = 0; a < cnt; a += sub, ++g) { ctx->pk_lo[g] = lo + a; ctx->pk_hi[g] = lo + (a + sub < cnt ? a + sub : cnt); }
                ctx->pk_first[c + 1] = g;
            }
            ctx->pk_total = g;
            double acc = 0.0;  // spread the packed sub-ranges evenly (error diffusion)
            for (int k = 0; k < g; ++k) { acc += ctx->pack_frac; ctx->pk_packed[k] = acc >= 0.999999; if (ctx->pk_packed[k]) acc -= 1.0; }
        }
        unsigned long long *h_flags = const_cast<unsigned long long *>(ctx->h_seq);
        for (int c = 0; c < C; ++c) {
            const int lo = ctx->ch_lo[c], cnt = ctx->ch_lo[c + 1] - lo;
            if (split_advance) {
                rc = launch_range(ctx, d_act, ctx->step_out, adv, 0, s0, 0, c, lo, cnt);
                if (rc) return rc;
                tmark(ctx, "advance_end", c, s0);
                cudaEvent_t ea = ctx->ev_adv[c];
                CK(cudaEventRecord(ea, s0));
                CK(cudaStreamWaitEvent(s_obs, ea, 0));
                // pose, target, reward, status ... are final once k_advance has run: they travel now, not behind the Reeds-Shepp kernels
                CK(cudaStreamWaitEvent(s_copy, ea, 0));
                rc = copy_fields(ctx, h_out, BY_ADVANCE, s_copy, lo, cnt);
                if (rc) return rc;
                advance_copied = true;
            }
            rc = launch_range(ctx, d_act, ctx->step_out, side, 0, s_obs, 0, c, lo, cnt, nullptr, false);
            if (rc) return rc;
            last_chunk = c;
            tmark(ctx, "observe_end", c, s_obs);
            if (ctx->packing && (stages & HOPE_STAGE_OBSERVE) && !(ctx->host_debug == 2)) {  // keep the beams that differ from the no-hit constant
                unsigned *count = ctx->d_lcount + ctx->pk_first[c];
                CK(cudaMemsetAsync(count, 0, sizeof(unsigned) * (ctx->pk_first[c + 1] - ctx->pk_first[c]), s_obs));
                k_pack_lidar<<<(cnt * 32 + 127) / 128, 128, 0, s_obs>>>(cnt, ctx->step_out.lidar + (size_t)lo * NRAY, make_tables(ctx).lidar_base, ctx->par.lidar_range,
                                                                       ctx->d_lbits + 4 * (size_t)lo, ctx->d_loff + lo, ctx->d_lpacked + (size_t)lo * NRAY, count, ctx->pk_sub);
                ctx->launches++;
            }
            tmark(ctx, "pack_end", c, s_obs);
            cudaEvent_t ev = ctx->ev_chunk[c];
            CK(cudaEventRecord(ev, s_obs));
            CK(cudaStreamWaitEvent(s_copy, ev, 0));
            // Flag A behind the sub-range counts: the host can issue the kept-value copies.  Flag B behind the other narrow
            // arrays (step counts, beam flags, offsets): the host can expand this range.
            if (ctx->packing) {
                const int nsub = ctx->pk_first[c + 1] - ctx->pk_first[c];
                CK(cudaMemcpyAsync(ctx->h_lcount + ctx->pk_first[c], ctx->d_lcount + ctx->pk_first[c], sizeof(unsigned) * nsub, cudaMemcpyDeviceToHost, s_copy));
                CK(cudaMemcpyAsync(h_flags + c, ctx->d_seq, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_copy));
                CK(cudaMemcpyAsync(ctx->h_lbits + 4 * (size_t)lo, ctx->d_lbits + 4 * (size_t)lo, sizeof(uint32_t) * 4 * cnt, cudaMemcpyDeviceToHost, s_copy));
                CK(cudaMemcpyAsync(ctx->h_loff + lo, ctx->d_loff + lo, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, s_copy));
                ctx->io_d2h_static += (size_t)cnt * 20 + 4 * nsub + 8;
            }
            if (ctx->expanding) {
                CK(cudaMemcpyAsync(ctx->h_mask_steps + (size_t)lo * NACT, ctx->stage_out.mask_steps + (size_t)lo * NACT, (size_t)cnt * NACT,
                                   cudaMemcpyDeviceToHost, s_copy));
                ctx->io_d2h_static += (size_t)cnt * NACT;
            }
            if (flagging) {
                CK(cudaMemcpyAsync(h_flags + hope_ctx::MAX_CHUNK_EVENTS + c, ctx->d_seq, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_copy));
                ctx->io_d2h_static += 8;
            }
            tmark(ctx, "flags_copied", c, s_copy);
            rc = copy_fields(ctx, h_out, BY_OBSERVE, s_copy, lo, cnt);
            if (rc) return rc;
            tmark(ctx, "observe_copies_end", c, s_copy);
        }
    }
    if (stages & HOPE_STAGE_RS) {
        // the persistent Reeds-Shepp grids would occupy every SM and starve the (later launched) k_observe ranges
        // whatever the stream priorities, so they start behind the last range and run under the copies instead
        if (side && ctx->host_rs_after_observe) CK(cudaStreamWaitEvent(s0, ctx->ev_chunk[last_chunk], 0));
        tmark(ctx, "rs_begin", -1, s0);
        rc = launch_range(ctx, d_act, ctx->step_out, HOPE_STAGE_RS, 0, s0, 0, 0, 0, n, nullptr, false);
        if (rc) return rc;
        tmark(ctx, "rs_end", -1, s0);
    }
    rc = copy_fields(ctx, h_out, side ? (advance_copied ? BY_RS : BY_ADVANCE | BY_RS) : BY_ANY, s0, 0, n);
    if (rc) return rc;
    tmark(ctx, "last_copies_end", -1, s0);
    if (side) {  // s_obs joins through the last range's event, s_copy joins here
        CK(cudaEventRecord(ctx->ev_join[1], s_copy));
        CK(cudaStreamWaitEvent(s0, ctx->ev_join[1], 0));
    }
    return HOPE_OK;
}

int hope_expand_mask(const uint8_t *h_steps, double *h_mask, int n) {
    if (!h_steps || !h_mask || n < 0) return HOPE_ERR_INVALID;
    hope_wire::expand_mask(h_steps, h_mask, 0, (size_t)n, 0);
    return HOPE_OK;
}

int hope_expand_lidar(const uint32_t *h_bits, const uint32_t *h_off, const double *h_packed, const double *h_nohit, double *h_lidar, int n, int portable) {
    if (!h_bits || !h_off || !h_packed || !h_nohit || !h_lidar || n < 0) return HOPE_ERR_INVALID;
    hope_wire::expand_lidar(h_bits, h_off, h_packed, h_nohit, h_lidar, 0, (size_t)n, portable);
    return HOPE_OK;
}

int hope_expand_mask_portable(const uint8_t *h_steps, double *h_mask, int n) {
    if (!h_steps || !h_mask || n < 0) return HOPE_ERR_INVALID;
    hope_wire::expand_mask(h_steps, h_mask, 0, (size_t)n, 1);
    return HOPE_OK;
}

int hope_host_wire_info(const hope_ctx *ctx, uint64_t info[8]) {
    if (!ctx || !info) return HOPE_ERR_INVALID;
    info[0] = ctx->io_h2d; info[1] = ctx->io_d2h; info[2] = ctx->expanding ? 1 : 0; info[3] = ctx->packing ? 1 : 0;
    info[4] = (uint64_t)ctx->host_threads; info[5] = (uint64_t)hope_wire::vector_path(); info[6] = (uint64_t)ctx->hm_chunks;
    info[7] = 0;
    for (int g = 0; g < ctx->pk_total; ++g) info[7] += ctx->packing && ctx->pk_packed[g] ? (uint64_t)(ctx->pk_hi[g] - ctx->pk_lo[g]) : 0;  // envs whose lidar travelled packed
    return HOPE_OK;
}

int hope_step_host(hope_ctx *ctx, const double *h_action, const hope_host_out *h_out, unsigned stages) {
    if (!ctx || !h_out) return HOPE_ERR_INVALID;  // h_action == NULL: CarParking.step(None), no motion
    if (!ctx->have_tables) return HOPE_ERR_NO_TABLES;
    if (!ctx->have_reset) return HOPE_ERR_NO_SCENES;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_stage(ctx, h_out->img != nullptr);
    if (rc) return rc;
    stages |= HOPE_STAGE_ADVANCE;
    // Software pipeline (see enqueue_host_step): k_observe runs env range by env range and each range's observation
    // buffers start their D2H copy behind it, under the Reeds-Shepp kernels.  Envs are independent, so the split
    // changes nothing.
    cudaStream_t s0 = ctx->lanes[0].main;
    rc = join_device_calls(ctx, s0);  // a device-API call may still be running on the caller's stream
    if (rc) return rc;
    ctx->dev_pending = false;         // everything below is ordered behind s0 and this call returns synchronised
    ++ctx->host_steps;
    ctx->tracing = ctx->host_trace > 0 && ctx->host_steps >= (unsigned long long)ctx->host_trace && (ctx->host_steps - ctx->host_trace) % 50 == 0;
    if (ctx->tracing) ctx->trace_t0 = std::chrono::steady_clock::now();
    if (ctx->host_graph_enabled && !ctx->profile && !ctx->tracing) {
        const bool same = ctx->host_graph && ctx->hg_action == h_action && ctx->hg_stages == stages &&
                          memcmp(&ctx->hg_out, h_out, sizeof(hope_out)) == 0;
        if (!same) {
            if (ctx->host_graph) { cudaGraphExecDestroy(ctx->host_graph); ctx->host_graph = nullptr; }
            plan_zero_copy(ctx, h_out);
            rc = plan_wire(ctx, h_out, stages);
            if (rc) return rc;
            const unsigned long long before = ctx->launches;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_host_step(ctx, h_action, h_out, stages);
            cudaError_t ce = cudaStreamEndCapture(s0, &g);
            if (rc == HOPE_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&ctx->host_graph, g, 0) == cudaSuccess) {
                ctx->hg_action = h_action; ctx->hg_out = *h_out; ctx->hg_stages = stages;
                ctx->hg_launches = ctx->launches - before;
                ctx->launches = before;  // nothing ran yet: the capture only recorded the launches
            } else {
                ctx->host_graph = nullptr; ctx->host_graph_enabled = false;  // fall back to direct enqueue for good
                ctx->launches = before;
                (void)cudaGetLastError();
            }
            if (g) cudaGraphDestroy(g);
        }
        if (ctx->host_graph) {
            CK(cudaGraphLaunch(ctx->host_graph, s0));
            ctx->launches += ctx->hg_launches;
            return finish_host_step(ctx, h_out, s0);
        }
    }
    plan_zero_copy(ctx, h_out);
    rc = plan_wire(ctx, h_out, stages);
    if (rc) return rc;
    rc = enqueue_host_step(ctx, h_action, h_out, stages);
    if (rc) return rc;
    hmark(ctx, "enqueued", -1);
    rc = finish_host_step(ctx, h_out, s0);
    trace_print(ctx);
    return rc;
}

int hope_reset_host(hope_ctx *ctx, const int32_t *h_scene_ids, const hope_host_out *h_out) {
    if (!ctx || !h_out) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_stage(ctx, h_out->img != nullptr);
    if (rc) return rc;
    hope_out reset_out = ctx->stage_out;
    if (!h_out->img) reset_out.img = nullptr;  // nobody reads the staged image: skip the render
    rc = join_device_calls(ctx, ctx->own_stream);
    if (rc) return rc;
    rc = hope_reset(ctx, h_scene_ids, &reset_out, ctx->own_stream);
    if (rc) return rc;
    const unsigned keep_mask = ctx->zero_copy_mask;
    ctx->zero_copy_mask = 0;  // the reset step always goes through the staging buffers
    rc = copy_fields(ctx, h_out, BY_ANY, ctx->own_stream, 0, ctx->n);
    ctx->zero_copy_mask = keep_mask;
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->own_stream));
    ctx->dev_pending = false;
    return HOPE_OK;
}

int hope_get_state(hope_ctx *ctx, double *h_pose, int32_t *h_t, double *h_accum, int32_t *h_scene_id) {
    if (!ctx) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    if (h_pose) CK(cudaMemcpy(h_pose, ctx->d_pose, sizeof(double) * 3 * ctx->n, cudaMemcpyDeviceToHost));
    if (h_t) CK(cudaMemcpy(h_t, ctx->d_t, sizeof(int) * ctx->n, cudaMemcpyDeviceToHost));
    if (h_accum) CK(cudaMemcpy(h_accum, ctx->d_accum, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost));
    if (h_scene_id) CK(cudaMemcpy(h_scene_id, ctx->d_scene, sizeof(int) * ctx->n, cudaMemcpyDeviceToHost));
    return HOPE_OK;
}

int hope_set_state(hope_ctx *ctx, const double *h_pose, const int32_t *h_t, const double *h_accum) {
    if (!ctx) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    if (h_pose) CK(cudaMemcpy(ctx->d_pose, h_pose, sizeof(double) * 3 * ctx->n, cudaMemcpyHostToDevice));
    if (h_t) CK(cudaMemcpy(ctx->d_t, h_t, sizeof(int) * ctx->n, cudaMemcpyHostToDevice));
    if (h_accum) CK(cudaMemcpy(ctx->d_accum, h_accum, sizeof(double) * ctx->n, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->d_pending, 0, ctx->n));
    return HOPE_OK;
}

int hope_planner_actions(hope_ctx *ctx, const double *d_policy_action, const hope_out *d_last_out, double *d_action_out,
                         uint8_t *d_executing, double step_ratio, void *stream) {
    if (!ctx || !d_policy_action || !d_last_out || !d_action_out || !(step_ratio > 0)) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const size_t N = ctx->n;
    uint8_t *u = ctx->d_plan_u8;
    PlanState ps{ctx->d_plan_rem, u, u + 5 * N, u + 6 * N, u + 7 * N, u + 8 * N};
    k_planner<<<(ctx->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(ctx->n, ps, d_policy_action, *d_last_out, d_action_out,
                                                                                     d_executing, step_ratio);
    ctx->launches++;
    CK(cudaGetLastError());
    return mark_device_call(ctx, static_cast<cudaStream_t>(stream));
}

int hope_wait_observed(hope_ctx *ctx, void *stream) {
    if (!ctx) return HOPE_ERR_INVALID;
    if (!ctx->obs_done_valid) return HOPE_ERR_INVALID;  // the last step had no observation stage, or ran in several env ranges
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ctx->ev_obs_done, 0));
    return HOPE_OK;
}

int hope_planner_reset(hope_ctx *ctx, void *stream) {
    if (!ctx) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_plan_u8, 0, 9 * (size_t)ctx->n, static_cast<cudaStream_t>(stream)));
    return HOPE_OK;
}

int hope_fp64_peak_tflops(int device, double *tflops) {
    if (!tflops) return HOPE_ERR_INVALID;
    hope_ctx *ctx = nullptr;
    CK(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double *buf = nullptr;
    CK(cudaMalloc(&buf, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, threads>>>(buf, 2000, 0.999999, 1e-9);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    return HOPE_OK;
}

int hope_profile_enable(hope_ctx *ctx, int on) {
    if (!ctx) return HOPE_ERR_INVALID;
    ctx->profile = on != 0;
    ctx->profile_serial = on == 2;
    return HOPE_OK;
}

int hope_profile_read(hope_ctx *ctx, double h_ms[8], uint64_t h_launches[8]) {
    if (!ctx || !h_ms || !h_launches) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    for (int k = 0; k < 8; ++k) {
        auto &ev = ctx->prof_events[k];
        double total = 0.0;
        for (size_t i = 0; i + 1 < ev.size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) total += ms;
        }
        h_ms[k] = total;
        h_launches[k] = ev.size() / 2;
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
    }
    return HOPE_OK;
}

int hope_get_counters(hope_ctx *ctx, uint64_t h_counters[8]) {
    if (!ctx || !h_counters) return HOPE_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    unsigned long long tmp[8];
    CK(cudaMemcpy(tmp, ctx->d_counters, sizeof(tmp), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) h_counters[k] = tmp[k];
    h_counters[5] = ctx->launches;
    return HOPE_OK;
}

#ifdef HOPE_STATS
// instrumented builds only (not declared in include/hope_b200.h): read and optionally clear the work counters
int hope_debug_stats(uint64_t h_stats[64], int reset) {
    unsigned long long tmp[64];
    if (cudaDeviceSynchronize() != cudaSuccess) return HOPE_ERR_CUDA;
    if (cudaMemcpyFromSymbol(tmp, hope::g_stats, sizeof(tmp)) != cudaSuccess) return HOPE_ERR_CUDA;
    if (h_stats) for (int k = 0; k < 64; ++k) h_stats[k] = tmp[k];
    if (reset) { memset(tmp, 0, sizeof(tmp)); if (cudaMemcpyToSymbol(hope::g_stats, tmp, sizeof(tmp)) != cudaSuccess) return HOPE_ERR_CUDA; }
    return HOPE_OK;
}
#endif

}  // extern "C"
