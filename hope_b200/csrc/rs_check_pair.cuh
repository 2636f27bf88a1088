// rs_check_pair.cuh — k_rs_check with TWO tried words per warp (experimental, -DHOPE_CHK_PAIR=1; off by default).
//
// Why: a tried word has 111 samples on average = 14 saved walker states, so a warp that owns one word keeps 14 of its
// 32 lanes busy, and 95 % of the words are condemned in their first round (profiles/r01_kernel_stats_w.json).  Here
// lanes 0-15 own one word and lanes 16-31 the next work item (often the same env's next word, but not necessarily:
// the check environment is per lane).  Lane j of a half owns saved states j and j + 16, eight samples each.  Both
// halves run the same instruction stream; every vote is a full-warp ballot of which each half reads its own 16 bits.
// A half is condemned by its first bad sample and idles until the other half is done.
//
// Verdicts are the same as chunk_is_bad's by construction (same samples, same tests, "any bad sample" per word); the
// degenerate trailing-zero words (reeds_shepp.py:501-505) are not handled here: the caller runs those through
// chunk_is_bad one word at a time.  tests/test_rs_check_host.py replays the reference's recorded verdicts through this
// code on the CPU warp emulation.  Needs rs_check.cuh; included inside namespace hope.
#pragma once

// obstacle edge p1-p2 against the four edges of the vehicle box (car_parking_base.py:477-526): exact bbox rejects, then the
// reference's untoleranced line-line solve
__device__ __forceinline__ bool edge_hits_box(double2 p1, double2 p2, const double *bx, const double *by, double vxmin, double vxmax,
                                              double vymin, double vymax) {
    if ((p1.x < vxmin && p2.x < vxmin) || (p1.x > vxmax && p2.x > vxmax) || (p1.y < vymin && p2.y < vymin) || (p1.y > vymax && p2.y > vymax))
        return false;
    const double oxmax = dmax(p1.x, p2.x), oxmin = dmin(p1.x, p2.x), oymax = dmax(p1.y, p2.y), oymin = dmin(p1.y, p2.y);
    const double dd = p2.y - p1.y, ee = p1.x - p2.x, ff = p1.y * p2.x - p1.x * p2.y;  // :504-506
    bool hit = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int q2 = (q + 1) & 3;
        const double vx1 = bx[q], vy1 = by[q], vx2 = bx[q2], vy2 = by[q2];
        if ((vx1 < oxmin && vx2 < oxmin) || (vx1 > oxmax && vx2 > oxmax) || (vy1 < oymin && vy2 < oymin) || (vy1 > oymax && vy2 > oymax)) continue;
        const double a = vy2 - vy1, b = vx1 - vx2, c = vy1 * vx2 - vx1 * vy2;  // :477-479
        const double det = a * ee - b * dd;                                    // :509
        if (det == 0.0) continue;
        double rx, ry;
        div_pair(b * ff - c * ee, c * dd - a * ff, det, rx, ry);               // :512-513
        const bool okx = !(rx > oxmax) && !(rx < oxmin) && !(rx > dmax(vx1, vx2)) && !(rx < dmin(vx1, vx2));
        const bool oky = !(ry > oymax) && !(ry < oymin) && !(ry > dmax(vy1, vy2)) && !(ry < dmin(vy1, vy2));
        if (okx && oky) hit = true;
    }
    return hit;
}

constexpr unsigned PAIR_LO = 0x0000ffffu, PAIR_HI = 0xffff0000u;

// One sample per lane.  `live` (the same in all 16 lanes of a half) goes false when the half's word is condemned.
// Called by the whole warp; all votes are full-warp ballots.
__device__ __forceinline__ void pair_samples_hit(const CheckEnv &E, const hope_params &par, bool valid, unsigned hmask, double lx, double ly,
                                                 double lyaw, bool &live) {
    double gx, gy, gyaw;
    sample_to_global(lx, ly, lyaw, E.cg, E.sg, E.q0x, E.q0y, E.q0h, gx, gy, gyaw);
    const bool out = valid && (gx < E.xmin || gx > E.xmax || gy < E.ymin || gy > E.ymax);  // car_parking_base.py:462-464
    unsigned vm = __ballot_sync(HOPE_FULL_MASK, valid);   // lanes that still have to be tested (kept warp-uniform below)
    unsigned hb = __ballot_sync(HOPE_FULL_MASK, out);
    if (hb & hmask) { live = false; valid = false; }
    if (hb & PAIR_LO) vm &= PAIR_HI;
    if (hb & PAIR_HI) vm &= PAIR_LO;
    if (!vm) return;
    double cth, sth, bx[4], by[4];
    sincos(gyaw, &sth, &cth);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // :468-471
        bx[q] = cth * par.box_x[q] - sth * par.box_y[q] + gx;
        by[q] = sth * par.box_x[q] + cth * par.box_y[q] + gy;
    }
    const double vxmin = dmin(dmin(bx[0], bx[1]), dmin(bx[2], bx[3])), vxmax = dmax(dmax(bx[0], bx[1]), dmax(bx[2], bx[3]));
    const double vymin = dmin(dmin(by[0], by[1]), dmin(by[2], by[3])), vymax = dmax(dmax(by[0], by[1]), dmax(by[2], by[3]));
    const int n_lo = __shfl_sync(HOPE_FULL_MASK, E.nobs, 0), n_hi = __shfl_sync(HOPE_FULL_MASK, E.nobs, 16);
    const int n_max = n_lo > n_hi ? n_lo : n_hi;
    for (int ob = 0; ob < n_max; ++ob) {  // the scene blocks are HOPE_MAX_OBS wide, so reading slot ob >= nobs is safe (and ignored)
        const double4 bb = ld_aabb(E.aabb + ob);
        // disjoint boxes cannot produce a hit (:518-526), so this reject is exact
        const bool enter = valid && ob < E.nobs && !(vxmax < bb.x || bb.y < vxmin || vymax < bb.z || bb.w < vymin);
        unsigned em = __ballot_sync(HOPE_FULL_MASK, enter);
        if (!em) continue;
        const int nv = E.nvp[ob];
        double2 p1 = __ldg(E.verts + ob * MAXV);
        for (int j = 0; j < MAXV; ++j) {
            const double2 p2 = __ldg(E.verts + ob * MAXV + ((j + 1 >= nv) ? 0 : j + 1));
            const bool hit = enter && valid && j < nv && edge_hits_box(p1, p2, bx, by, vxmin, vxmax, vymin, vymax);
            hb = __ballot_sync(HOPE_FULL_MASK, hit);      // the warp votes after every obstacle edge
            if (hb & hmask) { live = false; valid = false; }
            if (hb & PAIR_LO) { vm &= PAIR_HI; em &= PAIR_HI; }
            if (hb & PAIR_HI) { vm &= PAIR_LO; em &= PAIR_LO; }
            if (!em) break;
            p1 = p2;
        }
        if (!vm) return;
    }
}

// Lanes 0-15: word slot / environment of one work item, lanes 16-31: of another (`have` = false for a half without an
// item: it must still pass a readable slot and environment, e.g. the other half's).  Returns, per half, whether the
// word leaves the map or touches an obstacle.  Neither word may be a trailing-zero word (s.end_lx == 0.0).
__device__ bool pair_is_bad(WordSlot &s, const CheckEnv &E, const hope_params &par, int lane, bool have) {
    const int hl = lane & 15;
    const unsigned hmask = (lane & 16) ? PAIR_HI : PAIR_LO;
    bool live = have;
    int chunk_base = 0;
    for (;;) {
        uint8_t code0 = live ? s.st_code[hl] : RS_DONE, code1 = live ? s.st_code[hl + 16] : RS_DONE;
        double pd0 = s.st_pd[hl], pd1 = s.st_pd[hl + 16];
        for (int r = 0; r < RS_STRIDE; ++r) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {  // round r of states 0-15, then of states 16-31 (only words longer than 128 samples)
                uint8_t &code = p ? code1 : code0;
                double &pd = p ? pd1 : pd0;
                const bool valid = live && code != RS_DONE;
                if (!__any_sync(HOPE_FULL_MASK, valid)) continue;
                double lx = 0.0, ly = 0.0, lyaw = 0.0;
                if (valid && code != RS_ORIGIN) {
                    const int sgi = code & 0x7F;
                    rs_interp(pd, (int)((s.types >> (4 * sgi)) & 0xF), E.maxc, s.org[sgi], lx, ly, lyaw);
                }
                pair_samples_hit(E, par, valid, hmask, lx, ly, lyaw, live);
                if (valid) walker_next(s.len, s.n, E.step, code, pd);
            }
            if (!__any_sync(HOPE_FULL_MASK, live && (code0 != RS_DONE || code1 != RS_DONE))) break;
        }
        const bool more = live && s.total < 0;  // a word longer than one chunk (rare): lane 0 of the half walks on
        if (!__any_sync(HOPE_FULL_MASK, more)) break;
        chunk_base += RS_CHUNK;
        __syncwarp();
        if (more && hl == 0) walk_chunk(s, s.len, E.step, chunk_base);
        __syncwarp();
    }
    return have && !live;
}
