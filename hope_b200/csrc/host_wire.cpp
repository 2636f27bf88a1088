// host_wire.cpp — host half of hope_step_host's narrow wire format (plain C++, no CUDA).
//
// Two of the arrays a step returns are redundant as float64 and cross PCIe in a lossless narrow form instead:
//   * the action mask [N][42] is a function of its uint8 step counts (action_mask.py:182-183): the counts travel;
//   * a lidar beam that hits nothing within range reads exactly `lidar_range - lidar_base[ray]` (lidar_simulator.py:46,
//     134: clip to the range, subtract the vehicle's own extent along the beam), a constant per ray.  k_pack_lidar keeps
//     only the values whose bits differ from that constant: 120 flag bits + one offset per env + the kept doubles travel.
// The routines here rebuild the caller's float64 arrays bit for bit.  AVX-512 (expand-load / two-table permute) when the
// CPU has it, portable C++ otherwise; both produce the same bytes (tests/test_cabi_and_host.py runs each against numpy).
#include "host_wire.h"

#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define HOPE_WIRE_X86 1
#else
#define HOPE_WIRE_X86 0
#endif

namespace hope_wire {

namespace {

constexpr int NRAY = 120, NACT = 42, NITER = 10;

void mask_rows_zero_fix(const uint8_t *steps, double *mask, size_t lo, size_t hi) {
    // action_mask.py:183: an env whose 42 step counts are all 0 gets 0.01 everywhere
    for (size_t i = lo; i < hi; ++i) {
        const uint8_t *sp = steps + i * NACT;
        uint64_t w[5];
        uint16_t t;
        memcpy(w, sp, 40); memcpy(&t, sp + 40, 2);
        if ((w[0] | w[1] | w[2] | w[3] | w[4] | t) == 0) {
            double *mp = mask + i * NACT;
            for (int j = 0; j < NACT; ++j) mp[j] = 0.01;
        }
    }
}

void expand_mask_portable(const uint8_t *steps, double *mask, size_t lo, size_t hi) {
    double lut[NITER + 1];
    for (int k = 0; k <= NITER; ++k) lut[k] = (double)k / 10;  // the same division k_observe evaluates
    for (size_t e = lo * NACT; e < hi * NACT; ++e) mask[e] = lut[steps[e] <= NITER ? steps[e] : NITER];
    mask_rows_zero_fix(steps, mask, lo, hi);
}

void expand_lidar_portable(const uint32_t *bits, const uint32_t *off, const double *packed, const double *nohit, double *lidar, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
        const uint32_t *b = bits + 4 * i;
        const double *p = packed + off[i];
        double *o = lidar + NRAY * i;
        memcpy(o, nohit, sizeof(double) * NRAY);
        for (int q = 0; q < 4; ++q) {
            uint32_t w = b[q];
            double *oq = o + 32 * q;
            while (w) {
                oq[__builtin_ctz(w)] = *p++;
                w &= w - 1;
            }
        }
    }
}

#if HOPE_WIRE_X86
// Both routines write with non-temporal stores when the destination allows it (64-byte aligned vectors): the rows are
// written once and read by somebody else later, and a regular store would first read each line from memory (the expansion
// is bound by the host's memory bandwidth: 85 MB written per 65 536-env step).
__attribute__((target("avx512f"))) void expand_mask_avx512(const uint8_t *steps, double *mask, size_t lo, size_t hi) {
    double lut[16];
    for (int k = 0; k < 16; ++k) lut[k] = (double)(k <= NITER ? k : NITER) / 10;
    const __m512d lut_lo = _mm512_loadu_pd(lut), lut_hi = _mm512_loadu_pd(lut + 8);
    const __m512i top = _mm512_set1_epi64(NITER);
    size_t e = lo * NACT;
    const size_t end = hi * NACT;
    for (; e < end && (reinterpret_cast<uintptr_t>(mask + e) & 63); ++e) mask[e] = lut[steps[e] <= NITER ? steps[e] : NITER];
    for (; e + 8 <= end; e += 8) {
        const __m512i idx = _mm512_min_epu64(_mm512_cvtepu8_epi64(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(steps + e))), top);
        _mm512_stream_pd(mask + e, _mm512_permutex2var_pd(lut_lo, idx, lut_hi));  // bit 3 of the index picks the second table
    }
    for (; e < end; ++e) mask[e] = lut[steps[e] <= NITER ? steps[e] : NITER];
    _mm_sfence();
    mask_rows_zero_fix(steps, mask, lo, hi);
}

__attribute__((target("avx512f,popcnt"))) void expand_lidar_avx512(const uint32_t *bits, const uint32_t *off, const double *packed, const double *nohit, double *lidar,
                                                                    size_t lo, size_t hi) {
    __m512d nh[NRAY / 8];
    for (int g = 0; g < NRAY / 8; ++g) nh[g] = _mm512_loadu_pd(nohit + 8 * g);
    const bool aligned = (reinterpret_cast<uintptr_t>(lidar) & 63) == 0;  // rows are 960 bytes = 15 lines
    for (size_t i = lo; i < hi; ++i) {
        const uint32_t *b = bits + 4 * i;
        const double *p = packed + off[i];
        double *o = lidar + NRAY * i;
        // 8 beams per instruction: the next 8 kept values are loaded whole (hence the 64 readable bytes behind `packed`) and
        // expanded into the flagged lanes, the constant elsewhere
        if (aligned) {
#pragma GCC unroll 15
            for (int g = 0; g < NRAY / 8; ++g) {
                const __mmask8 m = (__mmask8)(b[g >> 2] >> (8 * (g & 3)));
                _mm512_stream_pd(o + 8 * g, _mm512_mask_expand_pd(nh[g], m, _mm512_loadu_pd(p)));
                p += __builtin_popcount((unsigned)m);
            }
        } else {
#pragma GCC unroll 15
            for (int g = 0; g < NRAY / 8; ++g) {
                const __mmask8 m = (__mmask8)(b[g >> 2] >> (8 * (g & 3)));
                _mm512_storeu_pd(o + 8 * g, _mm512_mask_expand_pd(nh[g], m, _mm512_loadu_pd(p)));
                p += __builtin_popcount((unsigned)m);
            }
        }
    }
    _mm_sfence();
}

bool have_avx512() {
    static const bool yes = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("popcnt");
    return yes;
}
#endif

}  // namespace

void expand_mask(const uint8_t *steps, double *mask, size_t lo, size_t hi, int force_portable) {
#if HOPE_WIRE_X86
    if (!force_portable && have_avx512()) return expand_mask_avx512(steps, mask, lo, hi);
#endif
    (void)force_portable;
    expand_mask_portable(steps, mask, lo, hi);
}

void expand_lidar(const uint32_t *bits, const uint32_t *off, const double *packed, const double *nohit, double *lidar, size_t lo, size_t hi, int force_portable) {
#if HOPE_WIRE_X86
    if (!force_portable && have_avx512()) return expand_lidar_avx512(bits, off, packed, nohit, lidar, lo, hi);
#endif
    (void)force_portable;
    expand_lidar_portable(bits, off, packed, nohit, lidar, lo, hi);
}

int vector_path() {
#if HOPE_WIRE_X86
    return have_avx512() ? 1 : 0;
#else
    return 0;
#endif
}

}  // namespace hope_wire
