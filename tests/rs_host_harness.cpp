// CPU harness for tests/test_rs_words_host.py: compiles the Reeds-Shepp word enumeration the k_rs_enumerate kernel
// runs (hope_b200/csrc/rs_words.cuh, on top of hope_device.cuh) with g++ and the host libm, so the reference's
// known answers (tests/golden/reeds_shepp.npz, recorded from the unmodified reeds_shepp.py) can be replayed through
// the product's own code without a GPU.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__
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
