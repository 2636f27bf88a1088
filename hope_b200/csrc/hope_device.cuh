// hope_device.cuh — float64 device helpers shared by the ParkingEnv kernels (sm_100a).
//
// Everything here is compiled with -fmad=false: the reference is numpy / CPython float64 where
// every product and sum is rounded separately (SURVEY.md §7 "fp64 and no FMA contraction").
// The only fused operations are the explicit __fma_rn calls inside the exact-arithmetic
// fallbacks, where FMA is used to recover rounding errors, not to skip them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define HOPE_FULL_MASK 0xffffffffu
#define HOPE_PI 3.141592653589793

namespace hope {

// min / max of finite doubles as one compare + select (fmin/fmax carry NaN handling that costs twice
// as many instructions in SASS; no operand here is ever NaN)
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------
// Exact orientation sign.  Ring-vs-ring `intersects` (car_parking_base.py:153-158) is a robust
// predicate in GEOS; a float64 determinant alone can flip it for near-degenerate contacts.
// Stage A: Shewchuk's static filter.  Stage B: the determinant expanded into six exact
// two-term products, summed with a grow-expansion (exact), sign of the leading component.
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    double bv = __dsub_rn(s, a), av = __dsub_rn(s, bv);
    e = __dadd_rn(__dsub_rn(a, av), __dsub_rn(b, bv));
}
__device__ __forceinline__ void two_prod(double a, double b, double &p, double &e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
__device__ __noinline__ int orient_exact(double ax, double ay, double bx, double by, double cx, double cy,
                                         unsigned long long *fallback_counter) {
    double t[12], e[12];
    two_prod(ax, by, t[0], t[1]);   two_prod(-ax, cy, t[2], t[3]);  two_prod(-cx, by, t[4], t[5]);
    two_prod(-ay, bx, t[6], t[7]);  two_prod(ay, cx, t[8], t[9]);   two_prod(cy, bx, t[10], t[11]);
    int n = 0;
    for (int k = 0; k < 12; ++k) {
        double q = t[k];
        for (int i = 0; i < n; ++i) { double s, r; two_sum(q, e[i], s, r); e[i] = r; q = s; }
        e[n++] = q;
    }
    if (fallback_counter) atomicAdd(fallback_counter, 1ull);
    for (int i = n - 1; i >= 0; --i)
        if (e[i] != 0.0) return e[i] > 0.0 ? 1 : -1;
    return 0;
}
__device__ __forceinline__ int orient(double ax, double ay, double bx, double by, double cx, double cy,
                                      unsigned long long *fallback_counter) {
    const double eps = 1.1102230246251565e-16;
    const double k = (3.0 + 16.0 * eps) * eps;
    double l = (ax - cx) * (by - cy), r = (ay - cy) * (bx - cx);
    double det = l - r, bound = k * (fabs(l) + fabs(r));
    if (det > bound) return 1;
    if (det < -bound) return -1;
    return orient_exact(ax, ay, bx, by, cx, cy, fallback_counter);
}
__device__ __forceinline__ bool within(double px, double py, double ax, double ay, double bx, double by) {
    return dmin(ax, bx) <= px && px <= dmax(ax, bx) && dmin(ay, by) <= py && py <= dmax(ay, by);
}
// Closed segments share a point (proper crossing, touch or collinear overlap).
__device__ __forceinline__ bool segments_touch(double p1x, double p1y, double p2x, double p2y, double q1x, double q1y,
                                               double q2x, double q2y, unsigned long long *fc) {
    if (dmax(p1x, p2x) < dmin(q1x, q2x) || dmax(q1x, q2x) < dmin(p1x, p2x)) return false;
    if (dmax(p1y, p2y) < dmin(q1y, q2y) || dmax(q1y, q2y) < dmin(p1y, p2y)) return false;
    int o1 = orient(p1x, p1y, p2x, p2y, q1x, q1y, fc), o2 = orient(p1x, p1y, p2x, p2y, q2x, q2y, fc);
    int o3 = orient(q1x, q1y, q2x, q2y, p1x, p1y, fc), o4 = orient(q1x, q1y, q2x, q2y, p2x, p2y, fc);
    if (o1 * o2 < 0 && o3 * o4 < 0) return true;
    if (o1 == 0 && within(q1x, q1y, p1x, p1y, p2x, p2y)) return true;
    if (o2 == 0 && within(q2x, q2y, p1x, p1y, p2x, p2y)) return true;
    if (o3 == 0 && within(p1x, p1y, q1x, q1y, q2x, q2y)) return true;
    if (o4 == 0 && within(p2x, p2y, q1x, q1y, q2x, q2y)) return true;
    return false;
}

// Area of (convex quad S) ∩ (convex quad C): Sutherland-Hodgman against C's four half-planes,
// then the shoelace sum.  C must be counter-clockwise (the host stores dest boxes that way).
// Polygon.intersection().area in car_parking_base.py:164-170, 216-220.
__device__ __noinline__ double quad_clip_area(const double *sx, const double *sy, const double *cx, const double *cy) {
    double ax[12], ay[12], ox[12], oy[12];
    int n = 4;
    for (int i = 0; i < 4; ++i) { ox[i] = sx[i]; oy[i] = sy[i]; }
    for (int i = 0; i < 4 && n > 0; ++i) {
        int j = (i + 1) & 3, m = n;
        double Ax = cx[i], Ay = cy[i], ex = cx[j] - Ax, ey = cy[j] - Ay;
        for (int k = 0; k < m; ++k) { ax[k] = ox[k]; ay[k] = oy[k]; }
        n = 0;
        for (int k = 0; k < m; ++k) {
            int k2 = (k + 1 == m) ? 0 : k + 1;
            double px = ax[k], py = ay[k], qx = ax[k2], qy = ay[k2];
            double sp = ex * (py - Ay) - ey * (px - Ax), sq = ex * (qy - Ay) - ey * (qx - Ax);
            if (sp >= 0) {
                ox[n] = px; oy[n] = py; ++n;
                if (sq < 0) { double t = sp / (sp - sq); ox[n] = px + t * (qx - px); oy[n] = py + t * (qy - py); ++n; }
            } else if (sq >= 0) {
                double t = sp / (sp - sq); ox[n] = px + t * (qx - px); oy[n] = py + t * (qy - py); ++n;
            }
        }
    }
    if (n < 3) return 0.0;
    double s = 0.0;
    for (int i = 0; i < n; ++i) { int j = (i + 1 == n) ? 0 : i + 1; s += ox[i] * oy[j] - ox[j] * oy[i]; }
    return fabs(s) * 0.5;
}

// vehicle.py:32-36 — corners rb, rf, lf, lb of the box at (x, y, h); a*x + b*y + xoff left to right.
__device__ __forceinline__ void vehicle_box(double x, double y, double c, double s, const double *bxl, const double *byl,
                                            double *bx, double *by) {
    double ms = -s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bx[i] = c * bxl[i] + ms * byl[i] + x;
        by[i] = s * bxl[i] + c * byl[i] + y;
    }
}

// CPython float modulo for a positive modulus (reeds_shepp.py:585 uses `theta % (2*pi)`).
__device__ __forceinline__ double py_mod_pos(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) { if (m < 0.0) m += b; } else m = 0.0;
    return m;
}
// reeds_shepp.py:581-592
__device__ __forceinline__ double rs_M(double th) {
    double p = py_mod_pos(th, 2.0 * HOPE_PI);
    if (p > HOPE_PI) p -= 2.0 * HOPE_PI;
    return p;
}
// reeds_shepp.py:561-568
__device__ __forceinline__ double pi_2_pi(double th) {
    while (th > HOPE_PI) th -= 2.0 * HOPE_PI;
    while (th < -HOPE_PI) th += 2.0 * HOPE_PI;
    return th;
}

}  // namespace hope
