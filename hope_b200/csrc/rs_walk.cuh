// rs_walk.cuh — sampling a Reeds-Shepp word (reeds_shepp.py:452-537, 46-49): the literal `pd += d` chain of
// generate_local_course, interpolate, and the local -> map transform.  Plain functions (no thread or memory-space
// dependence) so that tests/rs_host_harness.cpp can compile the very same code with g++; included by hope_kernels.cu
// inside namespace hope.
#pragma once

constexpr int RS_STRIDE = 8;                    // samples per lane in one chunk
constexpr int RS_CHUNK = 32 * RS_STRIDE;        // samples covered by one set of saved states
constexpr uint8_t RS_ORIGIN = 0xFE, RS_END = 0x80, RS_DONE = 0xFF;

struct __align__(16) WordSlot {                 // 560 bytes, the sampling plan of one tried word
    double len[HOPE_RS_MAX_SEG];                // normalised signed segment lengths
    double org[HOPE_RS_MAX_SEG][5];             // per segment: ox, oy, oyaw, cos(oyaw), sin(oyaw) (local frame)
    double st_pd[32];                           // saved walker state at sample RS_STRIDE*j of the chunk
    double end_lx;                              // local x of the final end point (trailing-zero rule)
    double resume_pd;                           // walker state at the start of the next chunk
    uint8_t st_code[32];
    uint32_t types;                             // 4 bits per segment
    int n;                                      // segments
    int total;                                  // samples in the word if the walk reached the end, else -1
    uint8_t resume_code, pad[3];
};
static_assert(sizeof(WordSlot) % 16 == 0, "WordSlot is staged with 16-byte asynchronous copies");

// One step of generate_local_course's sample sequence (:452-507).  (code, pd) is the sample just
// emitted; on return it is the next one.  code: RS_ORIGIN = path start, seg index = loop sample of
// that segment at arc parameter pd, RS_END|seg = the final end point, RS_DONE = no more samples.
template <class LenT>
__device__ __forceinline__ void walker_next(const LenT &len, int nseg, double step, uint8_t &code, double &pd) {
    int seg;
    double d;
    if (code == RS_ORIGIN) {
        seg = 0;
        d = len[0] > 0.0 ? step : -step;
        pd = d - 0.0;                                   // pd = d - ll with ll = 0.0 (:471-472, :486)
    } else if (code & RS_END) {
        code = RS_DONE;
        return;
    } else {
        seg = code;
        d = len[seg] > 0.0 ? step : -step;
        pd += d;                                        // :492
    }
    for (;;) {
        double l = len[seg];
        if (fabs(pd) <= fabs(l)) { code = (uint8_t)seg; return; }   // :488
        if (seg + 1 == nseg) { code = (uint8_t)(RS_END | seg); pd = l; return; }  // :496-498
        double ll = l - pd - d;                         // :494
        double ln = len[seg + 1];
        d = ln > 0.0 ? step : -step;                    // :475-478
        pd = (l * ln > 0) ? -d - ll : d - ll;           // :483-486
        ++seg;
    }
}

// interpolate (:510-537) from a segment origin (ox, oy, oyaw, cos oyaw, sin oyaw); returns the local-frame pose of the sample.
__device__ __forceinline__ void rs_interp(double p, int m, double maxc, double ox, double oy, double oyaw, double oc, double os, double &lx, double &ly,
                                          double &lyaw) {
    if (m == HOPE_RS_S) {
        lx = ox + p / maxc * oc;
        ly = oy + p / maxc * os;
        lyaw = oyaw;
    } else {
        double sl, cl;
        sincos(p, &sl, &cl);
        double ldx = sl / maxc, ldy = (1.0 - cl) / (m == HOPE_RS_L ? maxc : -maxc);
        double cy_ = oc, sy_ = -os;             // cos(-oyaw), sin(-oyaw)
        double gdx = cy_ * ldx + sy_ * ldy, gdy = -sy_ * ldx + cy_ * ldy;
        lx = ox + gdx; ly = oy + gdy;
        lyaw = (m == HOPE_RS_L) ? oyaw + p : oyaw - p;
    }
}
__device__ __forceinline__ void rs_interp(double p, int m, double maxc, const double *org, double &lx, double &ly, double &lyaw) {
    rs_interp(p, m, maxc, org[0], org[1], org[2], org[3], org[4], lx, ly, lyaw);
}

// Replay up to RS_CHUNK samples of a word's chain from its resume state, saving a state every
// RS_STRIDE.  Inside a segment the step is the bare `pd += d; |pd| <= |l|` of the reference;
// everything else (origin, segment changes, end point) goes through walker_next.
template <class LenT>
__device__ void walk_chunk(WordSlot &s, const LenT &len, double step, int chunk_base) {
    uint8_t code = s.resume_code;
    double pd = s.resume_pd;
    const int nseg = s.n;
    int k = 0, cur = -1;
    double d = 0.0, al = 0.0;
    if (code < HOPE_RS_MAX_SEG) { cur = code; const double l = len[code]; al = fabs(l); d = l > 0.0 ? step : -step; }
    while (k < RS_CHUNK && code != RS_DONE) {
        s.st_code[k / RS_STRIDE] = code; s.st_pd[k / RS_STRIDE] = pd;
        int i = 0;
        while (i < RS_STRIDE && code != RS_DONE) {
            if (code == cur) {
                while (i < RS_STRIDE) {
                    const double nx = pd + d;           // reeds_shepp.py:492
                    if (!(fabs(nx) <= al)) break;       // :488
                    pd = nx; ++i;
                }
                if (i == RS_STRIDE) break;
            }
            walker_next(len, nseg, step, code, pd);
            ++i;
            if (code < HOPE_RS_MAX_SEG) { cur = code; const double l = len[code]; al = fabs(l); d = l > 0.0 ? step : -step; }
            else cur = -1;
        }
        k += i;
    }
    for (int j = (k + RS_STRIDE - 1) / RS_STRIDE; j < 32; ++j) s.st_code[j] = RS_DONE;
    s.resume_code = code; s.resume_pd = pd;
    s.total = (code == RS_DONE) ? chunk_base + k : -1;
}

// The sampling plan of one tried word: segment origins in the local frame (each segment starts where the previous one
// ends, reeds_shepp.py:468-507 through interpolate) and the first chunk of the walk.
__device__ void plan_word(WordSlot &s, const RsWord &w, double maxc, double step) {
    // (the lengths stay a local array: a register-resident select chain for len[seg] measured 25 % slower on B200, the
    // local loads hit L1)
    double len[HOPE_RS_MAX_SEG];
    uint32_t ty = 0;
#pragma unroll
    for (int k = 0; k < HOPE_RS_MAX_SEG; ++k) { len[k] = w.len[k]; s.len[k] = w.len[k]; ty |= (uint32_t)(w.types[k] & 0xF) << (4 * k); }
    s.types = ty; s.n = w.n;
    double ox = 0.0, oy = 0.0, oyaw = 0.0, oc = 1.0, os = 0.0;  // scalars, not an array: sincos' output pointers would put an array on the stack
    for (int k = 0; k < w.n; ++k) {
        s.org[k][0] = ox; s.org[k][1] = oy; s.org[k][2] = oyaw; s.org[k][3] = oc; s.org[k][4] = os;
        double ex, ey, eyaw;
        rs_interp(len[k], (int)((ty >> (4 * k)) & 0xF), maxc, ox, oy, oyaw, oc, os, ex, ey, eyaw);  // end of segment k = origin of k+1
        ox = ex; oy = ey;
        if (eyaw != oyaw) { oyaw = eyaw; double sn, cs; sincos(eyaw, &sn, &cs); os = sn; oc = cs; }
    }
    s.end_lx = ox;
    s.resume_code = RS_ORIGIN; s.resume_pd = 0.0;
    walk_chunk(s, len, step, 0);
}

// calc_all_paths' transform of a local-frame sample into the map frame (reeds_shepp.py:46-49); cg, sg = cos(-q0h), sin(-q0h)
__device__ __forceinline__ void sample_to_global(double lx, double ly, double lyaw, double cg, double sg, double q0x, double q0y, double q0h,
                                                 double &gx, double &gy, double &gyaw) {
    gx = cg * lx + sg * ly + q0x; gy = -sg * lx + cg * ly + q0y;  // :47-48
    gyaw = pi_2_pi(lyaw + q0h);                                    // :49
}
