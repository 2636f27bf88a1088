// advance.cuh — per-warp scratch and the pooled ring-vs-ring collision test of k_advance.  The statements each thread of
// k_advance runs are in advance_body.inc, included textually into the kernel (hope_kernels.cu) and into
// tests/advance_host_harness.cpp, which compiles them with g++ on the CPU warp emulation (tests/warp_emu.h) and replays the
// pose / status / reward traces recorded from the unmodified reference.  (Textual fragments: the kernel's machine code stays
// byte-identical to the build verified on the GPU.)  Needs hope_types.cuh, hope_device.cuh, ld_aabb; included inside namespace hope.
#pragma once

// HOPE_ADV_STAGED: each lane keeps an outward-rounded float copy of its scene's obstacle bounding boxes in shared memory for the
// whole step (8 KB per warp at 16 rings), so the ten substeps' broad phase reads shared memory instead of waiting for a global
// load per obstacle per substep (19 % of k_advance's warp samples sat on that compare).  A float reject implies the exact
// reject (boxes rounded away from each other's interior); a float overlap is confirmed on the doubles.  The 128-ring build keeps
// the per-obstacle loads (64 KB per warp would not fit).
#ifndef HOPE_ADV_STAGED
#define HOPE_ADV_STAGED (HOPE_MAX_OBS <= 16)
#endif

struct AdvanceSmem {           // per warp
    double bx[32][4], by[32][4];   // current vehicle box of each lane's env
    int sid[32];
    uint16_t queue[32 * MAXO];     // (lane << 8) | obstacle of every vehicle-AABB / obstacle-AABB overlap
#if HOPE_ADV_STAGED
    float4 box[MAXO][32];          // [obstacle][lane]: xmin (rounded down), xmax (up), ymin (down), ymax (up); +inf/-inf = no obstacle
#endif
};

#if HOPE_ADV_STAGED
// once per step: this lane's obstacle boxes -> shared memory (independent loads, issued back to back)
__device__ __forceinline__ void stage_obstacle_boxes(AdvanceSmem &sm, const Pool &pool, int sid, int nobs, int lane) {
    const double4 *aabb = reinterpret_cast<const double4 *>(pool.aabb) + (size_t)sid * MAXO;
#pragma unroll
    for (int k = 0; k < MAXO; ++k) {
        float4 f = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
        if (k < nobs) {
            const double4 bb = ld_aabb(aabb + k);
            f = make_float4(__double2float_rd(bb.x), __double2float_ru(bb.y), __double2float_rd(bb.z), __double2float_ru(bb.w));
        }
        sm.box[k][lane] = f;
    }
}
#endif

// Per-lane result: does lane's vehicle ring touch any obstacle ring of its scene
// (car_parking_base.py:153-158)?  `check` selects the lanes that ask.  Phase 1: every asking lane
// walks its obstacle AABBs (exact reject) and enqueues the overlaps.  Phase 2: the warp drains the
// queue two items at a time, 16 lanes per item = 4 vehicle edges x up to 4 obstacle edges, one
// robust segment-pair test per lane.
__device__ __forceinline__ unsigned warp_collisions(bool check, const double *bx, const double *by, int sid, int nobs,
                                                    const Pool &pool, AdvanceSmem &sm, int lane, unsigned long long *fc) {
    if (!__any_sync(HOPE_FULL_MASK, check)) return 0u;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) { sm.bx[lane][i] = bx[i]; sm.by[lane][i] = by[i]; }
    sm.sid[lane] = sid;
    const double vxmin = dmin(dmin(bx[0], bx[1]), dmin(bx[2], bx[3])), vxmax = dmax(dmax(bx[0], bx[1]), dmax(bx[2], bx[3]));
    const double vymin = dmin(dmin(by[0], by[1]), dmin(by[2], by[3])), vymax = dmax(dmax(by[0], by[1]), dmax(by[2], by[3]));
    const double4 *aabb = reinterpret_cast<const double4 *>(pool.aabb) + (size_t)sid * MAXO;
    int qn = 0;
#if HOPE_ADV_STAGED
    {
        // broad phase on the staged float boxes, no warp collective inside; exact confirmation only for the float overlaps
        const float fxmin = __double2float_rd(vxmin), fxmax = __double2float_ru(vxmax), fymin = __double2float_rd(vymin), fymax = __double2float_ru(vymax);
        unsigned cand = 0;
#pragma unroll
        for (int k = 0; k < MAXO; ++k) {
            const float4 b = sm.box[k][lane];
            if (!(fxmax < b.x || b.y < fxmin || fymax < b.z || b.w < fymin)) cand |= 1u << k;
        }
        if (!check) cand = 0;
        unsigned over = 0;
        while (cand) {
            const int k = __ffs(cand) - 1;
            cand &= cand - 1;
            const double4 bb = ld_aabb(aabb + k);  // xmin xmax ymin ymax
            if (!(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin)) over |= 1u << k;  // disjoint boxes: exact reject
        }
        const int cnt = __popc(over);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_sync(HOPE_FULL_MASK, incl, (lane - o) & 31); if (lane >= o) incl += v; }
        qn = __shfl_sync(HOPE_FULL_MASK, incl, 31);
        int pos = incl - cnt;
        while (over) {
            const int k = __ffs(over) - 1;
            over &= over - 1;
            sm.queue[pos++] = (uint16_t)((lane << 8) | k);
        }
    }
#else
    int maxn = check ? nobs : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(HOPE_FULL_MASK, maxn, o));
    for (int k = 0; k < maxn; ++k) {
        bool over = false;
        if (check && k < nobs) {
            double4 bb = ld_aabb(aabb + k);  // xmin xmax ymin ymax
            over = !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin);  // disjoint boxes: exact reject
        }
        unsigned m = __ballot_sync(HOPE_FULL_MASK, over);
        if (over) sm.queue[qn + __popc(m & ((1u << lane) - 1))] = (uint16_t)((lane << 8) | k);
        qn += __popc(m);
    }
#endif
    __syncwarp();
    unsigned collided = 0;
    const int half = lane >> 4, pair = lane & 15, vi = pair & 3, oj = pair >> 2;
    for (int base = 0; base < qn; base += 2) {
        const int item = base + half;
        bool hit = false;
        if (item < qn) {
            const int code = sm.queue[item], owner = code >> 8, k = code & 255;
            if (!((collided >> owner) & 1)) {
                const int osid = sm.sid[owner];
                const int nv = pool.nv[(size_t)osid * MAXO + k];
                if (oj < nv) {
                    const double2 *v = reinterpret_cast<const double2 *>(pool.obs) + ((size_t)osid * MAXO + k) * MAXV;
                    const double2 p = __ldg(v + oj), q = __ldg(v + ((oj + 1 == nv) ? 0 : oj + 1));
                    const int vi2 = (vi + 1) & 3;
                    hit = segments_touch(sm.bx[owner][vi], sm.by[owner][vi], sm.bx[owner][vi2], sm.by[owner][vi2], p.x, p.y, q.x, q.y, fc);
                }
            }
        }
        const unsigned m = __ballot_sync(HOPE_FULL_MASK, hit);
        if (m & 0xffffu) collided |= 1u << (sm.queue[base] >> 8);
        if ((m >> 16) && base + 1 < qn) collided |= 1u << (sm.queue[base + 1] >> 8);
    }
    __syncwarp();
    return collided;
}

__device__ __forceinline__ double angle_gap(double a1, double a2) {  // car_parking_base.py:203-206
    double d = acos(cos(a1 - a2));
    return d < HOPE_PI / 2 ? d : HOPE_PI - d;
}
