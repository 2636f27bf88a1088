// rs_check_pooled.cuh — k_rs_check with the line-pair tests of a round pooled over the whole warp (the shipped variant,
// HOPE_CHK_POOLED = 1; measured on B200: step 9 % faster than the per-lane edge loop of rs_check.cuh, outputs bit-identical,
// profiles/r02_ab_variants.jsonl).
//
// Why: in the shipped loop a lane whose vehicle box overlaps an obstacle's bounding box walks that obstacle's edges on its
// own, and only 5.4 of 32 lanes are in that loop when any is (profiles/r01_kernel_stats_w.json) — the most expensive part
// of a round runs at a sixth of the warp's width.  Here, as in k_advance's collision test (advance.cuh), every lane first
// only enqueues its (sample, obstacle) overlaps; then the warp drains the queue two items at a time, 16 lanes per item =
// 4 obstacle edges x 4 vehicle edges, ONE line-pair test per lane, and votes after every pass.  The same pairs are tested
// with the same arithmetic, and a word is bad if any pair hits, so the verdicts are those of chunk_is_bad (rs_check.cuh).
// tests/test_rs_check_host.py and tests/test_step_host.py replay the reference's recorded verdicts through this code on the
// CPU warp emulation.  Needs rs_check.cuh; included inside namespace hope.
#pragma once

#ifndef HOPE_CHK_POOL_GROUP
#define HOPE_CHK_POOL_GROUP 3
#endif
constexpr int POOL_GROUP = HOPE_CHK_POOL_GROUP;  // obstacles enqueued before the queue is drained

struct CheckSmem {                  // per warp
    double bx[32][4], by[32][4];    // vehicle box of each lane's sample
    uint16_t queue[32 * POOL_GROUP];  // (lane << 8) | obstacle of every vehicle-box / obstacle-box overlap of the group
};

// obstacle edge p1-p2 against vehicle edge v1-v2 (car_parking_base.py:477-526): exact bbox rejects, then the reference's
// untoleranced line-line solve with both segments' bbox tests
__device__ __forceinline__ bool edge_pair_hit(double2 p1, double2 p2, double vx1, double vy1, double vx2, double vy2) {
    const double oxmax = dmax(p1.x, p2.x), oxmin = dmin(p1.x, p2.x), oymax = dmax(p1.y, p2.y), oymin = dmin(p1.y, p2.y);
    if ((vx1 < oxmin && vx2 < oxmin) || (vx1 > oxmax && vx2 > oxmax) || (vy1 < oymin && vy2 < oymin) || (vy1 > oymax && vy2 > oymax)) return false;
    const double dd = p2.y - p1.y, ee = p1.x - p2.x, ff = p1.y * p2.x - p1.x * p2.y;  // :504-506
    const double a = vy2 - vy1, b = vx1 - vx2, c = vy1 * vx2 - vx1 * vy2;              // :477-479
    const double det = a * ee - b * dd;                                                // :509
    if (det == 0.0) return false;
    double rx, ry;
    div_pair(b * ff - c * ee, c * dd - a * ff, det, rx, ry);                           // :512-513
    const bool okx = !(rx > oxmax) && !(rx < oxmin) && !(rx > dmax(vx1, vx2)) && !(rx < dmin(vx1, vx2));
    const bool oky = !(ry > oymax) && !(ry < oymin) && !(ry > dmax(vy1, vy2)) && !(ry < dmin(vy1, vy2));
    return okx && oky;
}

// Same contract as warp_samples_hit (rs_check.cuh): called by the whole warp, warp-uniform verdict, `mine` set for the lanes
// whose sample is bad (exact for all lanes only when `early` is false; with `early` the function returns at the first hit).
__device__ __forceinline__ bool pooled_samples_hit(const CheckEnv &E, const hope_params &par, bool valid, bool early, double lx, double ly,
                                                   double lyaw, bool &mine, CheckSmem &cs, int lane) {
    double gx, gy, gyaw;
    sample_to_global(lx, ly, lyaw, E.cg, E.sg, E.q0x, E.q0y, E.q0h, gx, gy, gyaw);
    mine = valid && (gx < E.xmin || gx > E.xmax || gy < E.ymin || gy > E.ymax);      // car_parking_base.py:462-464
    if (early && __any_sync(HOPE_FULL_MASK, mine)) return true;
    double cth, sth, bx[4], by[4];
    sincos(gyaw, &sth, &cth);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // :468-471
        bx[q] = cth * par.box_x[q] - sth * par.box_y[q] + gx;
        by[q] = sth * par.box_x[q] + cth * par.box_y[q] + gy;
        cs.bx[lane][q] = bx[q]; cs.by[lane][q] = by[q];
    }
    const double vxmin = dmin(dmin(bx[0], bx[1]), dmin(bx[2], bx[3])), vxmax = dmax(dmax(bx[0], bx[1]), dmax(bx[2], bx[3]));
    const double vymin = dmin(dmin(by[0], by[1]), dmin(by[2], by[3])), vymax = dmax(dmax(by[0], by[1]), dmax(by[2], by[3]));
    // Obstacles are taken in groups of POOL_GROUP: enqueue the (sample, obstacle) pairs of the group whose boxes overlap
    // (disjoint boxes cannot produce a hit, :518-526), drain, next group.  98 % of the condemned words are condemned by one of
    // the first three obstacles of their scene, so the first group usually ends the round.
    unsigned bad_lanes = 0;
    const int half = lane >> 4, pr = lane & 15, vi = pr & 3, oj = pr >> 2;
    for (int ob0 = 0; ob0 < E.nobs; ob0 += POOL_GROUP) {  // same trip count in every lane
        const int ob1 = min(ob0 + POOL_GROUP, E.nobs);
        int qn = 0;
        for (int ob = ob0; ob < ob1; ++ob) {
            const double4 bb = ld_aabb(E.aabb + ob);
            const bool over = valid && !mine && !((bad_lanes >> lane) & 1) && !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin);
            const unsigned m = __ballot_sync(HOPE_FULL_MASK, over);
            if (over) cs.queue[qn + __popc(m & ((1u << lane) - 1))] = (uint16_t)((lane << 8) | ob);
            qn += __popc(m);
        }
        __syncwarp();
        // two queue items per pass, 16 lanes per item = 4 obstacle edges x 4 vehicle edges, one line pair per lane
        for (int base = 0; base < qn; base += 2) {
            const int item = base + half;
            bool hit = false;
            if (item < qn) {
                const int code = cs.queue[item], owner = code >> 8, ob = code & 255;
                if (!((bad_lanes >> owner) & 1)) {
                    const int nv = E.nvp[ob];
                    if (oj < nv) {
                        const double2 p1 = __ldg(E.verts + ob * MAXV + oj), p2 = __ldg(E.verts + ob * MAXV + ((oj + 1 == nv) ? 0 : oj + 1));
                        const int vi2 = (vi + 1) & 3;
                        hit = edge_pair_hit(p1, p2, cs.bx[owner][vi], cs.by[owner][vi], cs.bx[owner][vi2], cs.by[owner][vi2]);
                    }
                }
            }
            const unsigned m = __ballot_sync(HOPE_FULL_MASK, hit);
            if (early && m) { __syncwarp(); return true; }  // one bad sample condemns the word
            if (m & 0xffffu) bad_lanes |= 1u << (cs.queue[base] >> 8);
            if ((m >> 16) && base + 1 < qn) bad_lanes |= 1u << (cs.queue[base + 1] >> 8);
        }
        __syncwarp();  // the queue is rewritten by the next group
    }
    __syncwarp();  // the scratch is rewritten by the next sample
    if ((bad_lanes >> lane) & 1) mine = true;
    return __any_sync(HOPE_FULL_MASK, mine);
}

// chunk_is_bad (rs_check.cuh) with the pooled sample test
__device__ bool chunk_is_bad_pooled(const WordSlot &s, const CheckEnv &E, const hope_params &par, int lane, CheckSmem &cs) {
    uint8_t code = s.st_code[lane];
    double pd = s.st_pd[lane];
    const bool zero_tail = s.end_lx == 0.0;  // reeds_shepp.py:501-505, see chunk_is_bad
    unsigned zero_hits = 0, nonzero_bits = 0;
    for (int r = 0; r < RS_STRIDE; ++r) {
        const bool valid = code != RS_DONE;
        if (!__any_sync(HOPE_FULL_MASK, valid)) break;
        double lx = 0.0, ly = 0.0, lyaw = 0.0;
        if (valid && code != RS_ORIGIN) {
            const int sgi = code & 0x7F;
            rs_interp(pd, (int)((s.types >> (4 * sgi)) & 0xF), E.maxc, s.org[sgi], lx, ly, lyaw);
        }
        bool mine;
        const bool any_hit = pooled_samples_hit(E, par, valid, !zero_tail, lx, ly, lyaw, mine, cs, lane);
        if (!zero_tail) { if (any_hit) return true; }
        else if (valid) {  // degenerate goal: a hit on an x == 0.0 sample only counts if a later sample has x != 0.0
            if (lx != 0.0) { nonzero_bits |= 1u << r; }
            else if (mine) { zero_hits |= 1u << r; mine = false; }
        }
        if (zero_tail && __any_sync(HOPE_FULL_MASK, valid && mine)) return true;
        if (valid) walker_next(s.len, s.n, E.step, code, pd);
    }
    if (zero_tail) {  // lanes own consecutive sample ranges: scan from the last sample backwards
        bool nz_after = false, bad = false;
        for (int l2 = 31; l2 >= 0; --l2) {
            const unsigned nzb = __shfl_sync(HOPE_FULL_MASK, nonzero_bits, l2), zhb = __shfl_sync(HOPE_FULL_MASK, zero_hits, l2);
            for (int r = RS_STRIDE - 1; r >= 0; --r) {
                if (((zhb >> r) & 1) && nz_after) bad = true;
                if ((nzb >> r) & 1) nz_after = true;
            }
        }
        return bad;
    }
    return false;
}
