// rs_check.cuh — the warp-level part of k_rs_check (car_parking_base.py:452-534 over the samples of reeds_shepp.py:452-537):
// lane j resumes saved walker state j of a word's plan and evaluates its RS_STRIDE consecutive samples; the warp votes on
// "any sample bad?".  No thread-index or memory-space dependence beyond the warp intrinsics (__any_sync, __ballot_sync,
// __shfl_sync, __ldg), so tests/rs_check_host_harness.cpp compiles this very code with g++ on top of a 32-fiber warp
// emulation and checks its verdicts against is_traj_valid verdicts recorded from the unmodified reference.
// Needs hope_device.cuh, div_pair.cuh, rs_walk.cuh, ld_aabb, MAXV and HOPE_STAT; included by hope_kernels.cu inside
// namespace hope.
#pragma once

struct CheckEnv {
    double q0x, q0y, q0h, cg, sg, xmin, xmax, ymin, ymax, maxc, step;
    const double4 *aabb; const double2 *verts; const uint8_t *nvp; int nobs;
};

// Does ANY lane's vehicle box at local-frame pose (lx, ly, lyaw) leave the map or touch an obstacle edge?
// Called by the whole warp (lanes without a sample pass valid = false); returns a warp-uniform verdict and
// (with `early`) stops at the first obstacle for which some lane reports a hit, since one bad sample condemns the word.
// `mine` is set for the lanes that hit (only consulted by the degenerate trailing-zero rule).
__device__ __forceinline__ bool warp_samples_hit(const CheckEnv &E, const hope_params &par, bool valid, bool early, double lx, double ly, double lyaw, bool &mine) {
    double gx, gy, gyaw;
    sample_to_global(lx, ly, lyaw, E.cg, E.sg, E.q0x, E.q0y, E.q0h, gx, gy, gyaw);
    mine = valid && (gx < E.xmin || gx > E.xmax || gy < E.ymin || gy > E.ymax);      // car_parking_base.py:462-464
#ifdef HOPE_STATS
    if (early && __any_sync(HOPE_FULL_MASK, mine)) { if ((threadIdx.x & 31) == 0) HOPE_STAT(14, 1); return true; }
#endif
    if (early && __any_sync(HOPE_FULL_MASK, mine)) return true;
    double cth, sth, bx[4], by[4];
    sincos(gyaw, &sth, &cth);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // :468-471
        bx[q] = cth * par.box_x[q] - sth * par.box_y[q] + gx;
        by[q] = sth * par.box_x[q] + cth * par.box_y[q] + gy;
    }
    const double vxmin = dmin(dmin(bx[0], bx[1]), dmin(bx[2], bx[3])), vxmax = dmax(dmax(bx[0], bx[1]), dmax(bx[2], bx[3]));
    const double vymin = dmin(dmin(by[0], by[1]), dmin(by[2], by[3])), vymax = dmax(dmax(by[0], by[1]), dmax(by[2], by[3]));
#if HOPE_CHK_EDGE_EXIT
    for (int ob = 0; ob < E.nobs; ++ob) {  // same trip count in every lane
        double4 bb = ld_aabb(E.aabb + ob);
        // disjoint boxes cannot produce a hit (:518-526), so this reject is exact
        const bool enter = valid && !mine && !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin);
#ifdef HOPE_STATS
        {
            const unsigned mm = __ballot_sync(HOPE_FULL_MASK, enter);
            if ((threadIdx.x & 31) == 0) { HOPE_STAT(32, 1); HOPE_STAT(33, __popc(mm)); if (mm) HOPE_STAT(34, 1); }
        }
#endif
        if (!__any_sync(HOPE_FULL_MASK, enter)) continue;
        // The edge loop is warp-uniform (every lane looks at the same obstacle, so nv is the same) and the warp votes
        // after EACH obstacle edge: one bad sample condemns the word, and 95 % of the tried words are condemned in
        // their first round (profiles/r01_kernel_stats_w.json), so most of the remaining edge tests are never needed.
        const int nv = E.nvp[ob];
        double2 p1 = __ldg(E.verts + ob * MAXV);
        for (int j = 0; j < nv; ++j) {
            double2 p2 = __ldg(E.verts + ob * MAXV + ((j + 1 == nv) ? 0 : j + 1));
            if (enter && !mine &&
                !((p1.x < vxmin && p2.x < vxmin) || (p1.x > vxmax && p2.x > vxmax) ||
                  (p1.y < vymin && p2.y < vymin) || (p1.y > vymax && p2.y > vymax))) {
                const double oxmax = dmax(p1.x, p2.x), oxmin = dmin(p1.x, p2.x), oymax = dmax(p1.y, p2.y), oymin = dmin(p1.y, p2.y);
                const double dd = p2.y - p1.y, ee = p1.x - p2.x, ff = p1.y * p2.x - p1.x * p2.y;  // :504-506
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int q2 = (q + 1) & 3;
                    const double vx1 = bx[q], vy1 = by[q], vx2 = bx[q2], vy2 = by[q2];
                    if ((vx1 < oxmin && vx2 < oxmin) || (vx1 > oxmax && vx2 > oxmax) ||
                        (vy1 < oymin && vy2 < oymin) || (vy1 > oymax && vy2 > oymax)) continue;
                    const double a = vy2 - vy1, b = vx1 - vx2, c = vy1 * vx2 - vx1 * vy2;  // :477-479
                    const double det = a * ee - b * dd;                                    // :509
                    if (det == 0.0) continue;
                    double rx, ry;
                    div_pair(b * ff - c * ee, c * dd - a * ff, det, rx, ry);               // :512-513
                    const bool okx = !(rx > oxmax) && !(rx < oxmin) && !(rx > dmax(vx1, vx2)) && !(rx < dmin(vx1, vx2));
                    const bool oky = !(ry > oymax) && !(ry < oymin) && !(ry > dmax(vy1, vy2)) && !(ry < dmin(vy1, vy2));
                    if (okx && oky) mine = true;
                }
            }
#ifdef HOPE_STATS
            if (early && __any_sync(HOPE_FULL_MASK, mine)) {
                if ((threadIdx.x & 31) == 0) { HOPE_STAT(15, 1); HOPE_STAT(16 + (ob < 15 ? ob : 15), 1); HOPE_STAT(36 + (j < 3 ? j : 3), 1); }
                return true;
            }
#endif
            if (early && __any_sync(HOPE_FULL_MASK, mine)) return true;
            p1 = p2;
        }
    }
#else
    for (int ob = 0; ob < E.nobs; ++ob) {  // same trip count in every lane
        double4 bb = ld_aabb(E.aabb + ob);
#ifdef HOPE_STATS
        {
            const unsigned mm = __ballot_sync(HOPE_FULL_MASK, valid && !mine && !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin));
            if ((threadIdx.x & 31) == 0) { HOPE_STAT(32, 1); HOPE_STAT(33, __popc(mm)); if (mm) HOPE_STAT(34, 1); }
        }
#endif
        // a hit needs rx inside both segments' x-ranges and ry inside both y-ranges (:518-526): disjoint
        // boxes cannot produce one, so these rejects are exact
        if (valid && !mine && !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin)) {
            const int nv = E.nvp[ob];
            double2 p1 = __ldg(E.verts + ob * MAXV);
            for (int j = 0; j < nv && !mine; ++j) {
                double2 p2 = __ldg(E.verts + ob * MAXV + ((j + 1 == nv) ? 0 : j + 1));
                // obstacle edge box vs vehicle box, tested corner-wise (no min/max needed to reject)
                if (!((p1.x < vxmin && p2.x < vxmin) || (p1.x > vxmax && p2.x > vxmax) ||
                      (p1.y < vymin && p2.y < vymin) || (p1.y > vymax && p2.y > vymax))) {
                    const double oxmax = dmax(p1.x, p2.x), oxmin = dmin(p1.x, p2.x), oymax = dmax(p1.y, p2.y), oymin = dmin(p1.y, p2.y);
                    const double dd = p2.y - p1.y, ee = p1.x - p2.x, ff = p1.y * p2.x - p1.x * p2.y;  // :504-506
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int q2 = (q + 1) & 3;
                        const double vx1 = bx[q], vy1 = by[q], vx2 = bx[q2], vy2 = by[q2];
                        if ((vx1 < oxmin && vx2 < oxmin) || (vx1 > oxmax && vx2 > oxmax) ||
                            (vy1 < oymin && vy2 < oymin) || (vy1 > oymax && vy2 > oymax)) continue;
                        const double a = vy2 - vy1, b = vx1 - vx2, c = vy1 * vx2 - vx1 * vy2;  // :477-479
                        const double det = a * ee - b * dd;                                    // :509
                        if (det == 0.0) continue;
                        double rx, ry;
                        div_pair(b * ff - c * ee, c * dd - a * ff, det, rx, ry);               // :512-513
                        const bool okx = !(rx > oxmax) && !(rx < oxmin) && !(rx > dmax(vx1, vx2)) && !(rx < dmin(vx1, vx2));
                        const bool oky = !(ry > oymax) && !(ry < oymin) && !(ry > dmax(vy1, vy2)) && !(ry < dmin(vy1, vy2));
                        if (okx && oky) mine = true;
                    }
                }
                p1 = p2;
            }
        }
#ifdef HOPE_STATS
        if (early && __any_sync(HOPE_FULL_MASK, mine)) {
            if ((threadIdx.x & 31) == 0) { HOPE_STAT(15, 1); HOPE_STAT(16 + (ob < 15 ? ob : 15), 1); }
            return true;
        }
#endif
        if (early && __any_sync(HOPE_FULL_MASK, mine)) return true;
    }
#endif
    return __any_sync(HOPE_FULL_MASK, mine);
}

// Evaluate the up-to RS_STRIDE samples each lane owns in the current chunk of word slot `s`.
// Returns true (warp-uniform) if any sample leaves the map or touches an obstacle edge.
__device__ bool chunk_is_bad(const WordSlot &s, const CheckEnv &E, const hope_params &par, int lane) {
    uint8_t code = s.st_code[lane];
    double pd = s.st_pd[lane];
    // Trailing samples whose local x is exactly 0.0 are dropped by the reference (:501-505).  That can
    // only remove anything if the final end point itself has x == 0.0 (zero_tail), a degenerate goal.
    const bool zero_tail = s.end_lx == 0.0;
    unsigned zero_hits = 0, nonzero_bits = 0;
    for (int r = 0; r < RS_STRIDE; ++r) {
        const bool valid = code != RS_DONE;
        if (!__any_sync(HOPE_FULL_MASK, valid)) break;
#ifdef HOPE_STATS
        {
            const unsigned vm = __ballot_sync(HOPE_FULL_MASK, valid);
            if (lane == 0) { HOPE_STAT(12, __popc(vm)); HOPE_STAT(13, 1); }
        }
#endif
        double lx = 0.0, ly = 0.0, lyaw = 0.0;
        if (valid && code != RS_ORIGIN) {
            const int sgi = code & 0x7F;
            rs_interp(pd, (int)((s.types >> (4 * sgi)) & 0xF), E.maxc, s.org[sgi], lx, ly, lyaw);
        }
        bool mine;
        const bool any_hit = warp_samples_hit(E, par, valid, !zero_tail, lx, ly, lyaw, mine);
#ifdef HOPE_STATS
        if (!zero_tail && any_hit && lane == 0) HOPE_STAT(2 + (r < 8 ? r : 7), 1);
#endif
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
