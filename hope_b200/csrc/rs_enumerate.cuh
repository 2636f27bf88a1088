// rs_enumerate.cuh — find_rs_path's candidate list for one env (car_parking_base.py:413-444): the admitted words of
// rs_words.cuh in heapdict pop order, cut at 1.6 x the shortest once more than two were popped.  A header of its own so
// tests/rs_search_host_harness.cpp compiles the same code with g++; included by hope_kernels.cu inside namespace hope
// (after hope_types.cuh and rs_words.cuh).
#pragma once

__device__ int enumerate_env(int i, const Pool &pool, const EnvState &st, const Tables &tb, const RsScratch &rs, const hope_out &out) {
    // default outputs: no path (k_rs_select overwrites them for the env whose search succeeds)
    if (out.rs_found) out.rs_found[i] = 0;
    if (out.rs_nseg) out.rs_nseg[i] = 0;
    if (out.rs_L) out.rs_L[i] = 0.0;
    if (out.rs_ncand) out.rs_ncand[i] = 0;
    if (out.rs_ntried) out.rs_ntried[i] = 0;
    if (out.rs_types) for (int k = 0; k < 5; ++k) out.rs_types[5 * i + k] = HOPE_RS_NONE;
    if (out.rs_lengths) for (int k = 0; k < 5; ++k) out.rs_lengths[5 * i + k] = 0.0;
    if (!st.gate[i]) { rs.ntry[i] = 0; rs.ncand[i] = 0; return 0; }
    const double *meta = pool.meta + (size_t)st.scene[i] * META;
    WordList w;
    enumerate_words(st.pose[3 * i], st.pose[3 * i + 1], st.pose[3 * i + 2], meta[M_DEST], meta[M_DEST + 1], meta[M_DEST + 2], tb.maxc, w, st.counters);
    // find_rs_path (car_parking_base.py:431-444): heapdict pop order (priority-only binary heap:
    // sift-up stops at a strictly smaller parent, sift-down prefers left unless right is strictly
    // smaller), cut at the first word with L > 1.6 L_min once more than two were popped.
    int heap[MAXW], hn = 0;
    double Ls[MAXW];
    for (int k = 0; k < w.count; ++k) Ls[k] = w.L[k] / tb.maxc;  // reeds_shepp.py:52
    for (int k = 0; k < w.count; ++k) {
        int p = hn++;
        heap[p] = k;
        while (p > 0) {
            int up = (p - 1) >> 1;
            if (Ls[heap[up]] < Ls[heap[p]]) break;
            int tmp = heap[up]; heap[up] = heap[p]; heap[p] = tmp;
            p = up;
        }
    }
    RsWord *dst = rs.words + (size_t)i * MAXW;
    int ntry = 0, idx = 0;
    double lmin = -1.0;
    while (hn) {
        ++idx;
        int top = heap[0];
        --hn;
        if (hn) {
            heap[0] = heap[hn];
            int p = 0;
            for (;;) {
                int l = 2 * p + 1, r = 2 * p + 2, low = (l < hn && Ls[heap[l]] < Ls[heap[p]]) ? l : p;
                if (r < hn && Ls[heap[r]] < Ls[heap[low]]) low = r;
                if (low == p) break;
                int tmp = heap[low]; heap[low] = heap[p]; heap[p] = tmp;
                p = low;
            }
        }
        if (lmin < 0) lmin = Ls[top];
        if (Ls[top] > 1.6 * lmin && idx > 2) break;
        RsWord ww;
        for (int k = 0; k < 5; ++k) { ww.len[k] = w.len[top][k]; ww.types[k] = (uint8_t)((w.ty[top] >> (4 * k)) & 0xF); }
        ww.L = w.L[top]; ww.n = w.n[top]; ww.pad[0] = ww.pad[1] = 0;
        dst[ntry++] = ww;
    }
    rs.ntry[i] = (uint8_t)ntry;
    rs.ncand[i] = (uint8_t)w.count;
    if (out.rs_ncand) out.rs_ncand[i] = (uint8_t)w.count;
    return ntry;
}
