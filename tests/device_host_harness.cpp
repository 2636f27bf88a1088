// CPU harness for tests/test_device_helpers_host.py: compiles the float64 helpers of hope_b200/csrc/hope_device.cuh
// (exact orientation predicate, segment-touch test, convex clip area, Python-style angle wraps) with g++ so the code
// the kernels run can be fuzzed against the rational-arithmetic oracle (oracle/geom.py) on adversarial inputs.
// The CUDA intrinsics are given their IEEE meaning (the file is compiled with -ffp-contract=off, like -fmad=false).
#include <cmath>
#include <cstdint>

// <cuda_runtime.h> resolves to tests/host_stubs/cuda_runtime.h (empty)
#define __device__
#define __forceinline__ inline
#define __noinline__
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }

#include "../hope_b200/csrc/hope_device.cuh"

extern "C" {
int dev_orient(const double *p, unsigned long long *fallbacks) { return hope::orient(p[0], p[1], p[2], p[3], p[4], p[5], fallbacks); }
int dev_segments_touch(const double *s, unsigned long long *fallbacks) {
    return hope::segments_touch(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], fallbacks) ? 1 : 0;
}
double dev_quad_clip_area(const double *sx, const double *sy, const double *cx, const double *cy) { return hope::quad_clip_area(sx, sy, cx, cy); }
double dev_rs_M(double th) { return hope::rs_M(th); }
double dev_pi_2_pi(double th) { return hope::pi_2_pi(th); }
}
