// CPU harness for tests/test_rs_words_host.py: compiles the Reeds-Shepp word enumeration the k_rs_enumerate kernel
// runs (hope_b200/csrc/rs_words.cuh) and the sample walker of k_rs_walk / k_rs_check (rs_walk.cuh), on top of
// hope_device.cuh, with g++ and the host libm, so the reference's
// known answers (tests/golden/reeds_shepp.npz, recorded from the unmodified reeds_shepp.py) can be replayed through
// the product's own code without a GPU.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)
#define HOPE_CONSTANT static const
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

#include "../include/hope_b200.h"
#include "../hope_b200/csrc/hope_device.cuh"

namespace hope {
constexpr int MAXW = 16;  // as in hope_kernels.cu
#include "../hope_b200/csrc/rs_words.cuh"
#include "../hope_b200/csrc/rs_walk.cuh"
}  // namespace hope

// q = (sx, sy, syaw, gx, gy, gyaw).  Outputs in the reference's units: lengths and L divided by maxc (reeds_shepp.py:52).
extern "C" int rs_host_words(const double *q, double maxc, int *count, int *nseg, uint8_t *types /*[16][5]*/, double *lengths /*[16][5]*/,
                             double *L /*[16]*/, unsigned long long *counters /*[8]*/) {
    hope::WordList w;
    const bool room = hope::enumerate_words(q[0], q[1], q[2], q[3], q[4], q[5], maxc, w, counters);
    *count = w.count;
    for (int k = 0; k < w.count; ++k) {
        nseg[k] = w.n[k];
        L[k] = w.L[k] / maxc;
        for (int i = 0; i < 5; ++i) {
            const unsigned t = (w.ty[k] >> (4 * i)) & 0xF;
            types[5 * k + i] = t == 0xF ? 255 : (uint8_t)t;
            lengths[5 * k + i] = i < w.n[k] ? w.len[k][i] / maxc : 0.0;
        }
    }
    return room ? 0 : 1;
}

// Samples of word k of the query, the way k_rs_walk plans them and k_rs_check's lanes evaluate them (lane j resumes
// saved state j and emits RS_STRIDE consecutive samples; further chunks are walked on).  Map-frame poses go to
// gx/gy/gyaw[cap], local x to lx[cap]; returns the number of samples or -1.
extern "C" int rs_host_samples(const double *q, double maxc, double rs_step, int k, int cap, double *gx, double *gy, double *gyaw, double *lx) {
    using namespace hope;
    WordList w;
    enumerate_words(q[0], q[1], q[2], q[3], q[4], q[5], maxc, w, nullptr);
    if (k < 0 || k >= w.count) return -1;
    RsWord rw;
    std::memset(&rw, 0, sizeof(rw));
    for (int i = 0; i < 5; ++i) { rw.len[i] = w.len[k][i]; rw.types[i] = (uint8_t)((w.ty[k] >> (4 * i)) & 0xF); }
    rw.L = w.L[k]; rw.n = w.n[k];
    const double step = rs_step * maxc;
    static WordSlot s;
    plan_word(s, rw, maxc, step);
    double sh, ch;
    sincos(q[2], &sh, &ch);
    const double cg = ch, sg = -sh;  // cos(-q0h), sin(-q0h)
    int chunk_base = 0, n = 0;
    for (;;) {
        for (int lane = 0; lane < 32; ++lane) {
            uint8_t code = s.st_code[lane];
            double pd = s.st_pd[lane];
            for (int r = 0; r < RS_STRIDE && code != RS_DONE; ++r) {
                double x = 0.0, y = 0.0, yaw = 0.0;
                if (code != RS_ORIGIN) {
                    const int sgi = code & 0x7F;
                    rs_interp(pd, (int)((s.types >> (4 * sgi)) & 0xF), maxc, s.org[sgi], x, y, yaw);
                }
                const int idx = chunk_base + lane * RS_STRIDE + r;
                if (idx >= cap) return -1;
                lx[idx] = x;
                sample_to_global(x, y, yaw, cg, sg, q[0], q[1], q[2], gx[idx], gy[idx], gyaw[idx]);
                n = idx + 1;
                walker_next(s.len, s.n, step, code, pd);
            }
        }
        if (s.total >= 0) break;
        chunk_base += RS_CHUNK;
        walk_chunk(s, s.len, step, chunk_base);
    }
    return s.total == n ? n : -2;
}
