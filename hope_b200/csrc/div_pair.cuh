// div_pair.cuh — the shared-reciprocal division of the raycast (k_observe) and the trajectory check (k_rs_check).
// A header of its own so that tests/rs_check_host_harness.cpp compiles the very same code with g++ (the model test of
// the rounding argument is tests/test_division_model.py); included by hope_kernels.cu inside namespace hope.
#pragma once

// n1/den and n2/den, each correctly rounded (== IEEE division), sharing one reciprocal: with r = RN(1/den)
// and q = RN(n r), the residual n - den q is exact in an FMA and q + r (n - den q) rounds to RN(n/den)
// (Markstein).  Used where the reference divides two numerators by the same determinant.
__device__ __forceinline__ void div_pair(double n1, double n2, double den, double &q1, double &q2) {
    const double r = __drcp_rn(den);
    const double a = __dmul_rn(n1, r), b = __dmul_rn(n2, r);
    q1 = __fma_rn(__fma_rn(-den, a, n1), r, a);
    q2 = __fma_rn(__fma_rn(-den, b, n2), r, b);
}
