/* TEST INFRASTRUCTURE ONLY — scalar float64 CPU restatement of the reference ParkingEnv step.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (hope_b200/) never does.
 *
 * Follows (paths relative to /root/reference/src):
 *   env/vehicle.py:21-96            pose, box, Euler bicycle integrator
 *   env/car_parking_base.py:153-299 substep loop, collision, arrival, status, reward
 *   env/car_parking_base.py:372-381 target representation
 *   env/car_parking_base.py:413-534 Reeds-Shepp search and trajectory validity
 *   env/lidar_simulator.py:31-135   120-beam raycast
 *   model/action_mask.py:166-196    action-mask sweep + post-processing
 *   env/reeds_shepp.py              word enumeration and sampling
 *   env/env_wrapper.py:10-50        action rescale, reward shaping
 *
 * Pinning: validated against tests/golden/ (traces recorded from the unmodified reference run
 * under oracle/refshim; see oracle/make_golden.py).  The shapely/GEOS predicates are restated
 * (oracle/geom.py); parity with real GEOS is UNPINNED (no reference tests exist).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -fopenmp (oracle/Makefile).
 * No FMA contraction anywhere: the reference is numpy/CPython float64 with separate roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef MAX_OBS
#define MAX_OBS 16 /* -DMAX_OBS=128 builds the variant for Dragon Lake Parking scenes */
#endif
#define MAX_V 4
#define N_RAY 120
#define N_UP 1200
#define N_ACT 42
#define N_ITER 10
#define MAX_PATHS 48
#define PI 3.141592653589793

enum { ST_CONTINUE = 1, ST_ARRIVED = 2, ST_COLLIDED = 3, ST_OUTBOUND = 4, ST_OUTTIME = 5 };

/* configs.py:13-38 */
static const double WHEEL_BASE = 2.8;
static const double BOX_X[4] = {-0.93, 0.96 + 2.8, 0.96 + 2.8, -0.93};
static const double BOX_Y[4] = {-1.94 / 2, -1.94 / 2, 1.94 / 2, 1.94 / 2};

typedef struct {
    const double *ray_a;      /* [120]  sin(theta_i)   lidar_simulator.py:86-88 */
    const double *ray_b;      /* [120] -cos(theta_i) */
    const double *lidar_base; /* [120] LidarSimlator.vehicle_boundary  lidar_simulator.py:48-53 */
    const double *mask_base;  /* [120] ActionMask.vehicle_lidar_base   action_mask.py:21-29 */
    const double *dist_star;  /* [1200][42][10]  action_mask.py:114-143 */
    const double *w_lo;       /* [10] 1 - r/10   action_mask.py:161 */
    const double *w_hi;       /* [10] r/10 */
    double maxc;              /* tan(0.75)/2.8   car_parking_base.py:422 */
} orc_tables;

/* ------------------------------------------------------------------------------------------ */
/* exact orientation predicate (restates GEOS' robust orientation index)                       */
static inline void two_sum(double a, double b, double *s, double *e) {
    double x = a + b, bv = x - a, av = x - bv;
    *s = x; *e = (a - av) + (b - bv);
}
static inline void two_prod(double a, double b, double *p, double *e) {
    *p = a * b; *e = fma(a, b, -*p);
}
static int orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    /* (ax-cx)(by-cy)-(ay-cy)(bx-cx) = ax*by - ax*cy - cx*by - ay*bx + ay*cx + cy*bx  (cx*cy cancels) */
    double t[12], e[13];
    int n = 0, i, k;
    two_prod(ax, by, &t[0], &t[1]);  two_prod(-ax, cy, &t[2], &t[3]);  two_prod(-cx, by, &t[4], &t[5]);
    two_prod(-ay, bx, &t[6], &t[7]); two_prod(ay, cx, &t[8], &t[9]);   two_prod(cy, bx, &t[10], &t[11]);
    for (k = 0; k < 12; ++k) { /* grow-expansion: e stays non-overlapping, sum stays exact */
        double q = t[k];
        for (i = 0; i < n; ++i) { double s, r; two_sum(q, e[i], &s, &r); e[i] = r; q = s; }
        e[n++] = q;
    }
    for (i = n - 1; i >= 0; --i) if (e[i] != 0.0) return e[i] > 0 ? 1 : -1;
    return 0;
}
static int orient(double ax, double ay, double bx, double by, double cx, double cy) {
    const double eps = 1.1102230246251565e-16, bound_k = (3.0 + 16.0 * eps) * eps;
    double l = (ax - cx) * (by - cy), r = (ay - cy) * (bx - cx), det = l - r;
    double bound = bound_k * (fabs(l) + fabs(r));
    if (det > bound) return 1;
    if (det < -bound) return -1;
    return orient_exact(ax, ay, bx, by, cx, cy);
}
static inline int in_box(double px, double py, double ax, double ay, double bx, double by) {
    return fmin(ax, bx) <= px && px <= fmax(ax, bx) && fmin(ay, by) <= py && py <= fmax(ay, by);
}
static int seg_hit(double p1x, double p1y, double p2x, double p2y, double q1x, double q1y, double q2x, double q2y) {
    int o1, o2, o3, o4;
    if (fmax(p1x, p2x) < fmin(q1x, q2x) || fmax(q1x, q2x) < fmin(p1x, p2x)) return 0;
    if (fmax(p1y, p2y) < fmin(q1y, q2y) || fmax(q1y, q2y) < fmin(p1y, p2y)) return 0;
    o1 = orient(p1x, p1y, p2x, p2y, q1x, q1y); o2 = orient(p1x, p1y, p2x, p2y, q2x, q2y);
    o3 = orient(q1x, q1y, q2x, q2y, p1x, p1y); o4 = orient(q1x, q1y, q2x, q2y, p2x, p2y);
    if (o1 * o2 < 0 && o3 * o4 < 0) return 1;
    if (o1 == 0 && in_box(q1x, q1y, p1x, p1y, p2x, p2y)) return 1;
    if (o2 == 0 && in_box(q2x, q2y, p1x, p1y, p2x, p2y)) return 1;
    if (o3 == 0 && in_box(p1x, p1y, q1x, q1y, q2x, q2y)) return 1;
    if (o4 == 0 && in_box(p2x, p2y, q1x, q1y, q2x, q2y)) return 1;
    return 0;
}

/* vehicle.py:32-36 — affine_transform(VehicleBox, [c,-s,s,c,x,y]): a*x + b*y + xoff left to right */
static void make_box(const double pose[3], double bx[4], double by[4]) {
    double c = cos(pose[2]), s = sin(pose[2]), ms = -s;
    for (int i = 0; i < 4; ++i) {
        bx[i] = c * BOX_X[i] + ms * BOX_Y[i] + pose[0];
        by[i] = s * BOX_X[i] + c * BOX_Y[i] + pose[1];
    }
}

typedef struct {
    const double *obs; /* [MAX_OBS][MAX_V][2] */
    const int *nverts; /* [MAX_OBS], 0 = unused slot */
} scene_obs;

/* car_parking_base.py:153-158 — ring boundary vs ring boundary */
static int collides(const double bx[4], const double by[4], scene_obs so) {
    for (int k = 0; k < MAX_OBS; ++k) {
        int nv = so.nverts[k];
        const double *v = so.obs + k * MAX_V * 2;
        for (int j = 0; j < nv; ++j) {
            int j2 = (j + 1) % nv;
            for (int i = 0; i < 4; ++i) {
                int i2 = (i + 1) & 3;
                if (seg_hit(bx[i], by[i], bx[i2], by[i2], v[2 * j], v[2 * j + 1], v[2 * j2], v[2 * j2 + 1])) return 1;
            }
        }
    }
    return 0;
}

/* Polygon.intersection().area for two convex quads: Sutherland-Hodgman + shoelace (oracle/geom.py) */
static double shoelace(const double *px, const double *py, int n) {
    double s = 0.0;
    if (n < 3) return 0.0;
    for (int i = 0; i < n; ++i) { int j = (i + 1) % n; s += px[i] * py[j] - px[j] * py[i]; }
    return fabs(s) * 0.5;
}
static double clip_area(const double sx[4], const double sy[4], const double cx_[4], const double cy_[4]) {
    double cx[4], cy[4], ax[16], ay[16], ox[16], oy[16];
    int n = 4, i, k;
    double sa = 0.0;
    for (i = 0; i < 4; ++i) { int j = (i + 1) & 3; sa += cx_[i] * cy_[j] - cx_[j] * cy_[i]; }
    sa = 0.5 * sa;
    for (i = 0; i < 4; ++i) { int s = sa < 0 ? 3 - i : i; cx[i] = cx_[s]; cy[i] = cy_[s]; }
    for (i = 0; i < 4; ++i) { ox[i] = sx[i]; oy[i] = sy[i]; }
    for (i = 0; i < 4 && n > 0; ++i) {
        int j = (i + 1) & 3, m = n;
        double Ax = cx[i], Ay = cy[i], ex = cx[j] - Ax, ey = cy[j] - Ay;
        memcpy(ax, ox, sizeof(double) * m); memcpy(ay, oy, sizeof(double) * m);
        n = 0;
        for (k = 0; k < m; ++k) {
            int k2 = (k + 1) % m;
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
    return shoelace(ox, oy, n);
}

/* vehicle.py:69-96 with step_time = 1: 20 Euler mini-iterations, literal operation order */
static void ks_substep(double pose[3], double steer, double speed) {
    double v = fmin(fmax(speed, -2.5), 2.5), phi = fmin(fmax(steer, -0.75), 0.75);
    double x = pose[0], y = pose[1], h = pose[2];
    for (int i = 0; i < 20; ++i) {
        x += v * cos(h) * 0.05 / 20;
        y += v * sin(h) * 0.05 / 20;
        h += v * tan(phi) / WHEEL_BASE * 0.05 / 20;
    }
    pose[0] = x; pose[1] = y; pose[2] = h;
}

/* lidar_simulator.py:31-135 */
static void lidar(const double pose[3], scene_obs so, const orc_tables *tb, double out[N_RAY]) {
    double a = cos(pose[2]), b = sin(pose[2]);
    double xoff = -pose[0] * a - pose[1] * b, yoff = pose[0] * b - pose[1] * a, mb = -b;
    double best[N_RAY];
    int any = 0;
    for (int i = 0; i < N_RAY; ++i) best[i] = INFINITY;
    for (int k = 0; k < MAX_OBS; ++k) {
        int nv = so.nverts[k];
        double rx[MAX_V], ry[MAX_V];
        const double *v = so.obs + k * MAX_V * 2;
        if (!nv) continue;
        for (int j = 0; j < nv; ++j) { /* affine [a, b, -b, a, xoff, yoff] */
            rx[j] = a * v[2 * j] + b * v[2 * j + 1] + xoff;
            ry[j] = mb * v[2 * j] + a * v[2 * j + 1] + yoff;
        }
        /* the reference drops obstacles whose ring is >= 10 m away (:69); every hit on such a ring is
           >= 10 m and clips to 10, so keeping them cannot change the result */
        any = 1;
        for (int j = 0; j < nv; ++j) {
            int j2 = (j + 1) % nv;
            double x1 = rx[j], y1 = ry[j], x2 = rx[j2], y2 = ry[j2];
            double d = y2 - y1, e = x1 - x2, f = y1 * x2 - x1 * y2;
            double xmax = fmax(x1, x2), xmin = fmin(x1, x2), ymax = fmax(y1, y2), ymin = fmin(y1, y2);
            for (int i = 0; i < N_RAY; ++i) {
                double A = tb->ray_a[i], B = tb->ray_b[i];
                double det = A * e - B * d;
                int par = (det == 0);
                if (par) det = 1;
                double px = (B * f - 0 * e) / det, py = (0 * d - A * f) / det;
                if ((i < 30 || i >= 90) && px < -1e-8) px = 100;
                if (i >= 30 && i < 90 && px > 1e-8) px = 100;
                if (i < 60 && py < -1e-8) py = 100;
                if (i >= 60 && py > 1e-8) py = 100;
                if (px > xmax) px = 100;
                if (px < xmin) px = 100;
                if (py > ymax) py = 100;
                if (py < ymin) py = 100;
                if (par) px = 100;
                double r = sqrt(px * px + py * py);
                if (r < best[i]) best[i] = r;
            }
        }
    }
    for (int i = 0; i < N_RAY; ++i) {
        double r = any ? fmin(fmax(best[i], 0.0), 10.0) : 10.0;
        out[i] = r - tb->lidar_base[i];
    }
}

/* action_mask.py:166-196. steps_out: integer collision-free step count per action after the
   min-filter (0..10); mask_out = steps/10, or all 0.01 when every entry is 0. */
static void action_mask(const double lidar_obs[N_RAY], const orc_tables *tb, double mask_out[N_ACT], int steps_out[N_ACT]) {
    double L[N_RAY], d[N_UP];
    long s[N_ACT], f[N_ACT];
    for (int i = 0; i < N_RAY; ++i) L[i] = fmin(fmax(lidar_obs[i], 0.0), 10.0) + tb->mask_base[i];
    for (int j = 0; j < N_UP; ++j) {
        int q = j / 10, r = j % 10;
        d[j] = L[q] * tb->w_lo[r] + L[(q + 1) % N_RAY] * tb->w_hi[r];
    }
    for (int a = 0; a < N_ACT; ++a) s[a] = N_ITER;
    for (int j = 0; j < N_UP; ++j)
        for (int a = 0; a < N_ACT; ++a) {
            const double *row = tb->dist_star + ((size_t)j * N_ACT + a) * N_ITER;
            int k = 0;
            while (k < N_ITER && row[k] <= d[j]) ++k;
            if (k < s[a]) s[a] = k;
        }
    for (int half = 0; half < 2; ++half) {
        long *h = s + half * 21, *g = f + half * 21;
        h[0] -= 1; h[20] -= 1;
        for (int i = 0; i < 21; ++i) { /* 5-tap min, scipy 'reflect' == window truncated at the ends */
            long m = h[i];
            for (int t = -2; t <= 2; ++t) { int u = i + t; if (u < 0) u = -u - 1; if (u > 20) u = 41 - u; if (h[u] < m) m = h[u]; }
            g[i] = m;
        }
    }
    long tot = 0;
    for (int a = 0; a < N_ACT; ++a) { long c = f[a] < 0 ? 0 : (f[a] > 10 ? 10 : f[a]); steps_out[a] = (int)c; tot += c; }
    for (int a = 0; a < N_ACT; ++a) mask_out[a] = tot == 0 ? 0.01 : (double)steps_out[a] / 10;
}

/* math.hypot as CPython >= 3.10 computes it (Modules/mathmodule.c, vector_norm): power-of-two
   scaling, double-length squares, one differential correction.  It is NOT libm hypot and can
   differ from it in the last ulp; the recorded golden traces carry this version's bits. */
static double py_hypot(double a, double b) {
    double v[2] = {fabs(a), fabs(b)}, mx = fmax(v[0], v[1]);
    double csum = 1.0, frac1 = 0.0, frac2 = 0.0, x, h, z, zz, sh, sl, scale;
    int e;
    if (isinf(v[0]) || isinf(v[1])) return INFINITY;
    if (isnan(v[0]) || isnan(v[1])) return NAN;
    if (mx == 0.0) return mx;
    frexp(mx, &e);
    if (e < -1023) { return 2.2250738585072014e-308 * py_hypot(a / 2.2250738585072014e-308, b / 2.2250738585072014e-308); }
    scale = ldexp(1.0, -e);
    for (int i = 0; i < 2; ++i) {
        x = v[i] * scale;
        z = x * x; zz = fma(x, x, -z);
        sh = csum + z; sl = (csum - sh) + z;
        csum = sh; frac1 += zz; frac2 += sl;
    }
    h = sqrt(csum - 1.0 + (frac1 + frac2));
    z = -h * h; zz = fma(-h, h, -z);
    sh = csum + z; sl = (csum - sh) + z;
    csum = sh; frac1 += zz; frac2 += sl;
    x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);
    return h / scale;
}
#define hypot py_hypot

/* ------------------------------------------------------------------------------------------ */
/* Reeds-Shepp (reeds_shepp.py)                                                                */
static double py_mod(double a, double b) { /* CPython float_rem for b > 0 */
    double m = fmod(a, b);
    if (m != 0.0) { if (m < 0) m += b; } else m = copysign(0.0, b);
    return m;
}
static double Mwrap(double th) { /* :581-592 */
    double p = py_mod(th, 2.0 * PI);
    if (p < -PI) p += 2.0 * PI;
    if (p > PI) p -= 2.0 * PI;
    return p;
}
static double pi_2_pi(double th) { /* :561-568 */
    while (th > PI) th -= 2.0 * PI;
    while (th < -PI) th += 2.0 * PI;
    return th;
}
static int f_SLS(double x, double y, double phi, double *t, double *u, double *v) { /* :133-149 */
    phi = Mwrap(phi);
    if (y > 0.0 && 0.0 < phi && phi < PI * 0.99) {
        double xd = -y / tan(phi) + x;
        *t = xd - tan(phi / 2.0); *u = phi;
        *v = sqrt(pow(x - xd, 2.0) + pow(y, 2.0)) - tan(phi / 2.0);
        return 1;
    } else if (y < 0.0 && 0.0 < phi && phi < PI * 0.99) {
        double xd = -y / tan(phi) + x;
        *t = xd - tan(phi / 2.0); *u = phi;
        *v = -sqrt(pow(x - xd, 2.0) + pow(y, 2.0)) - tan(phi / 2.0);
        return 1;
    }
    return 0;
}
static int f_LSL(double x, double y, double phi, double *t, double *u, double *v) { /* :79-87 */
    double ax = x - sin(phi), ay = y - 1.0 + cos(phi);
    double uu = hypot(ax, ay), tt = atan2(ay, ax);
    if (tt >= 0.0) { double vv = Mwrap(phi - tt); if (vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; } }
    return 0;
}
static int f_LSR(double x, double y, double phi, double *t, double *u, double *v) { /* :90-103 */
    double ax = x + sin(phi), ay = y - 1.0 - cos(phi);
    double u1 = hypot(ax, ay), t1 = atan2(ay, ax);
    u1 = pow(u1, 2.0);
    if (u1 >= 4.0) {
        double uu = sqrt(u1 - 4.0), th = atan2(2.0, uu), tt = Mwrap(t1 + th), vv = Mwrap(tt - phi);
        if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
    return 0;
}
static int f_LRL(double x, double y, double phi, double *t, double *u, double *v) { /* :106-117 */
    double ax = x - sin(phi), ay = y - 1.0 + cos(phi);
    double u1 = hypot(ax, ay), t1 = atan2(ay, ax);
    if (u1 <= 4.0) {
        double uu = -2.0 * asin(0.25 * u1), tt = Mwrap(t1 + 0.5 * uu + PI), vv = Mwrap(phi - tt + uu);
        if (tt >= 0.0 && uu <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
    return 0;
}
static void tau_omega(double u, double v, double xi, double eta, double phi, double *tau, double *omega) { /* :228-243 */
    double delta = Mwrap(u - v), A = sin(u) - sin(delta), B = cos(u) - cos(delta) - 1.0;
    double t1 = atan2(eta * A - xi * B, xi * A + eta * B);
    double t2 = 2.0 * (cos(delta) - cos(v) - cos(u)) + 3.0;
    *tau = t2 < 0 ? Mwrap(t1 + PI) : Mwrap(t1);
    *omega = Mwrap(*tau - u + v - phi);
}
static int f_LRLRn(double x, double y, double phi, double *t, double *u, double *v) { /* :246-257 */
    double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
    if (rho <= 1.0) {
        double uu = acos(rho), tt, vv;
        tau_omega(uu, -uu, xi, eta, phi, &tt, &vv);
        if (tt >= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
    return 0;
}
static int f_LRLRp(double x, double y, double phi, double *t, double *u, double *v) { /* :260-272 */
    double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho = (20.0 - xi * xi - eta * eta) / 16.0;
    if (0.0 <= rho && rho <= 1.0) {
        double uu = -acos(rho);
        if (uu >= -0.5 * PI) {
            double tt, vv;
            tau_omega(uu, uu, xi, eta, phi, &tt, &vv);
            if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
        }
    }
    return 0;
}
static int f_LRSR(double x, double y, double phi, double *t, double *u, double *v) { /* :311-323 */
    double xi = x + sin(phi), eta = y - 1.0 - cos(phi);
    double rho = hypot(-eta, xi), theta = atan2(xi, -eta);
    if (rho >= 2.0) {
        double tt = theta, uu = 2.0 - rho, vv = Mwrap(tt + 0.5 * PI - phi);
        if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
    return 0;
}
static int f_LRSL(double x, double y, double phi, double *t, double *u, double *v) { /* :326-339 */
    double xi = x - sin(phi), eta = y - 1.0 + cos(phi);
    double rho = hypot(xi, eta), theta = atan2(eta, xi);
    if (rho >= 2.0) {
        double r = sqrt(rho * rho - 4.0), uu = 2.0 - r, tt = Mwrap(theta + atan2(r, -2.0)), vv = Mwrap(phi - 0.5 * PI - tt);
        if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
    return 0;
}
static int f_LRSLR(double x, double y, double phi, double *t, double *u, double *v) { /* :414-429 */
    double xi = x + sin(phi), eta = y - 1.0 - cos(phi);
    double rho = hypot(xi, eta);
    if (rho >= 2.0) {
        double uu = 4.0 - sqrt(rho * rho - 4.0);
        if (uu <= 0.0) {
            double tt = Mwrap(atan2((4.0 - uu) * xi - 2.0 * eta, -2.0 * xi + (uu - 4.0) * eta)), vv = Mwrap(tt - phi);
            if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
        }
    }
    return 0;
}

typedef int (*word_fn)(double, double, double, double *, double *, double *);
/* output layouts: how (t,u,v) fill the length vector */
enum { PAT_TUV, PAT_VUT, PAT_TUmUV, PAT_TUUV, PAT_THUV, PAT_VUHT, PAT_THUHV };
enum { S_ = 0, L_ = 1, R_ = 2 };
typedef struct { word_fn fn; int back; int pat; int n; uint8_t types[5]; } word_family;
/* generation order of generate_path (:549-555); each family is evaluated on the 4 reflections
   (x,y,phi) (-x,y,-phi) (x,-y,-phi) (-x,-y,phi) except SCS which uses only the 1st and 3rd */
static const word_family FAMILIES[] = {
    {f_LSL, 0, PAT_TUV, 3, {L_, S_, L_}},            /* CSC :153-167 */
    {f_LSR, 0, PAT_TUV, 3, {L_, S_, R_}},            /*     :169-183 */
    {f_LRL, 0, PAT_TUV, 3, {L_, R_, L_}},            /* CCC :189-203 */
    {f_LRL, 1, PAT_VUT, 3, {L_, R_, L_}},            /*     :209-223 */
    {f_LRLRn, 0, PAT_TUmUV, 4, {L_, R_, L_, R_}},    /* CCCC :276-290 */
    {f_LRLRp, 0, PAT_TUUV, 4, {L_, R_, L_, R_}},     /*      :292-306 */
    {f_LRSL, 0, PAT_THUV, 4, {L_, R_, S_, L_}},      /* CCSC :343-357 */
    {f_LRSR, 0, PAT_THUV, 4, {L_, R_, S_, R_}},      /*      :359-373 */
    {f_LRSL, 1, PAT_VUHT, 4, {L_, S_, R_, L_}},      /*      :379-393 */
    {f_LRSR, 1, PAT_VUHT, 4, {R_, S_, R_, L_}},      /*      :395-409 */
    {f_LRSLR, 0, PAT_THUHV, 5, {L_, R_, S_, L_, R_}},/* CCSCC :433-447 */
};

typedef struct { int n; unsigned npmask; uint8_t types[5]; double len[5]; double L; } rs_word;

/* builtin sum() as the recorded run executed it.  CPython >= 3.12 (what recorded tests/golden)
   adds exact `float` items with Neumaier compensation, but drops to plain left-to-right adds
   from the first item that is not an exact float.  In the env the ego heading is an np.float64
   (vehicle.py:93), so every word length computed by bare arithmetic on `phi` is an np.float64
   too (mask bit set below) while lengths that come out of math.* calls are exact floats.
   CPython 3.8 (the reference's README) adds left to right throughout; orc_set_py_sum(0) selects
   that.  The variants differ only in the last ulp of a word's total length L. */
static int g_py312_sum = 1;
void orc_set_py_sum(int compensated) { g_py312_sum = compensated; }
static double py_sum(const double *v, int n, unsigned npmask) {
    double hi = 0.0, lo = 0.0;
    int i = 0;
    if (g_py312_sum)
        for (; i < n && !((npmask >> i) & 1); ++i) {
            double x = v[i], t = hi + x;
            if (fabs(hi) >= fabs(x)) lo += (hi - t) + x; else lo += (x - t) + hi;
            hi = t;
        }
    if (lo != 0.0 && isfinite(lo)) hi += lo;
    for (; i < n; ++i) hi = hi + v[i];
    return hi;
}
static int add_word(rs_word *ws, int nw, int n, const uint8_t *types, const double *len, unsigned npmask, int *err) { /* set_path :57-76 */
    double tmp[5];
    for (int k = 0; k < nw; ++k) {
        if (ws[k].n != n || memcmp(ws[k].types, types, n)) continue;
        for (int i = 0; i < n; ++i) tmp[i] = ws[k].len[i] - len[i];
        if (py_sum(tmp, n, ws[k].npmask | npmask) <= 0.01) return nw;
    }
    for (int i = 0; i < n; ++i) tmp[i] = fabs(len[i]);
    double L = py_sum(tmp, n, npmask);
    if (L >= 1000.0) return nw;
    if (!(L >= 0.001)) { *err = 1; return nw; } /* the reference asserts here (:73) */
    ws[nw].n = n; ws[nw].L = L; ws[nw].npmask = npmask;
    memcpy(ws[nw].types, types, 5); memcpy(ws[nw].len, len, sizeof(double) * 5);
    return nw + 1;
}
static int enumerate_words(const double q0[3], const double q1[3], double maxc, rs_word *ws, int *err) { /* generate_path :540-557 */
    double dx = q1[0] - q0[0], dy = q1[1] - q0[1], dth = q1[2] - q0[2];
    double c = cos(q0[2]), s = sin(q0[2]);
    double x = (c * dx + s * dy) * maxc, y = (-s * dx + c * dy) * maxc, phi = dth;
    double t, u, v, len[5];
    uint8_t ty[5];
    int nw = 0;
    static const uint8_t SLS_T[5] = {S_, L_, S_}, SRS_T[5] = {S_, R_, S_};
    if (f_SLS(x, y, phi, &t, &u, &v)) { len[0] = t; len[1] = u; len[2] = v; len[3] = len[4] = 0; nw = add_word(ws, nw, 3, SLS_T, len, 2u, err); }
    if (f_SLS(x, -y, -phi, &t, &u, &v)) { len[0] = t; len[1] = u; len[2] = v; len[3] = len[4] = 0; nw = add_word(ws, nw, 3, SRS_T, len, 2u, err); }
    double xb = x * cos(phi) + y * sin(phi), yb = x * sin(phi) - y * cos(phi);
    for (unsigned fi = 0; fi < sizeof(FAMILIES) / sizeof(FAMILIES[0]); ++fi) {
        const word_family *F = &FAMILIES[fi];
        double X = F->back ? xb : x, Y = F->back ? yb : y;
        for (int r = 0; r < 4; ++r) {
            double sx = (r & 1) ? -X : X, sy = (r & 2) ? -Y : Y, sp = (r == 1 || r == 2) ? -phi : phi;
            if (!F->fn(sx, sy, sp, &t, &u, &v)) continue;
            double H = -0.5 * PI;
            memset(len, 0, sizeof len);
            switch (F->pat) {
            case PAT_TUV: len[0] = t; len[1] = u; len[2] = v; break;
            case PAT_VUT: len[0] = v; len[1] = u; len[2] = t; break;
            case PAT_TUmUV: len[0] = t; len[1] = u; len[2] = -u; len[3] = v; break;
            case PAT_TUUV: len[0] = t; len[1] = u; len[2] = u; len[3] = v; break;
            case PAT_THUV: len[0] = t; len[1] = H; len[2] = u; len[3] = v; break;
            case PAT_VUHT: len[0] = v; len[1] = u; len[2] = H; len[3] = t; break;
            case PAT_THUHV: len[0] = t; len[1] = H; len[2] = u; len[3] = H; len[4] = v; break;
            }
            if (r & 1) for (int i = 0; i < F->n; ++i) len[i] = -len[i];
            for (int i = 0; i < 5; ++i) { uint8_t b = i < F->n ? F->types[i] : 255; ty[i] = (r & 2) && b != S_ && b != 255 ? (uint8_t)(3 - b) : b; }
            static const unsigned NP_SLOT[] = {4u, 1u, 8u, 8u, 8u, 1u, 16u}; /* where v lands, per PAT_* */
            nw = add_word(ws, nw, F->n, ty, len, NP_SLOT[F->pat], err);
        }
    }
    return nw;
}

typedef struct { double *x, *y, *yaw; int cap; } traj_buf;
static void interp(int ind, double l, int m, double maxc, double ox, double oy, double oyaw, traj_buf *b) { /* :510-537 */
    if (m == S_) {
        b->x[ind] = ox + l / maxc * cos(oyaw);
        b->y[ind] = oy + l / maxc * sin(oyaw);
        b->yaw[ind] = oyaw;
    } else {
        double ldx = sin(l) / maxc, ldy = m == L_ ? (1.0 - cos(l)) / maxc : (1.0 - cos(l)) / (-maxc);
        double gdx = cos(-oyaw) * ldx + sin(-oyaw) * ldy, gdy = -sin(-oyaw) * ldx + cos(-oyaw) * ldy;
        b->x[ind] = ox + gdx; b->y[ind] = oy + gdy;
        b->yaw[ind] = m == L_ ? oyaw + l : oyaw - l;
    }
}
/* generate_local_course :452-507 on the normalised word, then the global transform of
   calc_all_paths :46-49.  Returns the sample count T. */
static int sample_word(const rs_word *w, double maxc, double step, const double q0[3], traj_buf *b) {
    int point_num = (int)(w->L / step) + w->n + 3;
    if (point_num > b->cap) {
        b->cap = point_num * 2;
        b->x = realloc(b->x, sizeof(double) * b->cap); b->y = realloc(b->y, sizeof(double) * b->cap);
        b->yaw = realloc(b->yaw, sizeof(double) * b->cap);
    }
    for (int i = 0; i < point_num; ++i) b->x[i] = b->y[i] = b->yaw[i] = 0.0;
    int ind = 1;
    double d = w->len[0] > 0.0 ? step : -step, pd = d, ll = 0.0;
    for (int i = 0; i < w->n; ++i) {
        double l = w->len[i];
        int m = w->types[i];
        d = l > 0.0 ? step : -step;
        double ox = b->x[ind], oy = b->y[ind], oyaw = b->yaw[ind];
        ind -= 1;
        if (i >= 1 && (w->len[i - 1] * w->len[i]) > 0) pd = -d - ll; else pd = d - ll;
        while (fabs(pd) <= fabs(l)) { ind += 1; interp(ind, pd, m, maxc, ox, oy, oyaw, b); pd += d; }
        ll = l - pd - d;
        ind += 1;
        interp(ind, l, m, maxc, ox, oy, oyaw, b);
    }
    int T = point_num;
    while (T > 0 && b->x[T - 1] == 0.0) --T;
    double cg = cos(-q0[2]), sg = sin(-q0[2]);
    for (int i = 0; i < T; ++i) {
        double ix = b->x[i], iy = b->y[i];
        b->x[i] = cg * ix + sg * iy + q0[0];
        b->y[i] = -sg * ix + cg * iy + q0[1];
        b->yaw[i] = pi_2_pi(b->yaw[i] + q0[2]);
    }
    return T;
}

/* car_parking_base.py:452-534 */
static int traj_valid(const traj_buf *b, int T, const double bounds[4], scene_obs so) {
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int i = 0; i < T; ++i) { mnx = fmin(mnx, b->x[i]); mxx = fmax(mxx, b->x[i]); mny = fmin(mny, b->y[i]); mxy = fmax(mxy, b->y[i]); }
    if (mnx < bounds[0] || mxx > bounds[1] || mny < bounds[2] || mxy > bounds[3]) return 0;
    double *vx = malloc(sizeof(double) * T * 4 * 2), *vy = vx + T * 4;
    double x_max = -INFINITY, x_min = INFINITY, y_max = -INFINITY, y_min = INFINITY;
    for (int i = 0; i < T; ++i) {
        double c = cos(b->yaw[i]), s = sin(b->yaw[i]);
        for (int k = 0; k < 4; ++k) {
            double X = c * BOX_X[k] - s * BOX_Y[k] + b->x[i], Y = s * BOX_X[k] + c * BOX_Y[k] + b->y[i];
            vx[i * 4 + k] = X; vy[i * 4 + k] = Y;
            x_max = fmax(x_max, X); x_min = fmin(x_min, X); y_max = fmax(y_max, Y); y_min = fmin(y_min, Y);
        }
    }
    x_max += 5; x_min -= 5; y_max += 5; y_min -= 5;
    int hit = 0, kept = 0;
    for (int k = 0; k < MAX_OBS && !hit; ++k) {
        int nv = so.nverts[k];
        const double *v = so.obs + k * MAX_V * 2;
        if (!nv) continue;
        int ax = 1, bx = 1, ay = 1, by = 1;
        for (int j = 0; j < nv; ++j) {
            ax &= v[2 * j] > x_max; bx &= v[2 * j] < x_min; ay &= v[2 * j + 1] > y_max; by &= v[2 * j + 1] < y_min;
        }
        if (ax || bx || ay || by) continue;
        kept = 1;
        for (int j = 0; j < nv && !hit; ++j) {
            int j2 = (j + 1) % nv;
            double x1 = v[2 * j], y1 = v[2 * j + 1], x2 = v[2 * j2], y2 = v[2 * j2 + 1];
            double d = y2 - y1, e = x1 - x2, f = y1 * x2 - x1 * y2;
            double oxmax = fmax(x1, x2), oxmin = fmin(x1, x2), oymax = fmax(y1, y2), oymin = fmin(y1, y2);
            for (int i = 0; i < T && !hit; ++i)
                for (int c4 = 0; c4 < 4; ++c4) {
                    int c5 = (c4 + 1) & 3;
                    double vx1 = vx[i * 4 + c4], vy1 = vy[i * 4 + c4], vx2 = vx[i * 4 + c5], vy2 = vy[i * 4 + c5];
                    double a = vy2 - vy1, bb = vx1 - vx2, c = vy1 * vx2 - vx1 * vy2;
                    double det = a * e - bb * d;
                    if (det == 0) continue;
                    double rx = (bb * f - c * e) / det, ry = (c * d - a * f) / det;
                    int okx = !(rx > oxmax) && !(rx < oxmin) && !(rx > fmax(vx1, vx2)) && !(rx < fmin(vx1, vx2));
                    int oky = !(ry > oymax) && !(ry < oymin) && !(ry > fmax(vy1, vy2)) && !(ry < fmin(vy1, vy2));
                    if (okx && oky) { hit = 1; break; }
                }
        }
    }
    free(vx);
    (void)kept;
    return !hit;
}

typedef struct { int found, nseg, ncand, ntried, T_last, err; uint8_t types[5]; double len[5]; double L; } rs_result;

/* find_rs_path :413-450, with heapdict's binary heap (oracle/refshim/heapdict.py) for the pop order */
static void find_rs_path(const double pose[3], const double dest[3], const double bounds[4], scene_obs so,
                         const orc_tables *tb, traj_buf *buf, rs_result *out) {
    rs_word ws[MAX_PATHS];
    int heap[MAX_PATHS], hn = 0, err = 0;
    memset(out, 0, sizeof *out);
    memset(out->types, 255, 5);
    int nw = enumerate_words(pose, dest, tb->maxc, ws, &err);
    out->ncand = nw; out->err = err;
    if (!nw) return;
    double Ls[MAX_PATHS];
    for (int k = 0; k < nw; ++k) Ls[k] = ws[k].L / tb->maxc;
    for (int k = 0; k < nw; ++k) { /* push: sift-up stops at a strictly smaller parent */
        int i = hn++;
        heap[i] = k;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (Ls[heap[p]] < Ls[heap[i]]) break;
            int tmp = heap[p]; heap[p] = heap[i]; heap[i] = tmp; i = p;
        }
    }
    double min_len = -1;
    int idx = 0;
    while (hn) {
        idx += 1;
        int top = heap[0];
        --hn;
        if (hn) {
            heap[0] = heap[hn];
            int i = 0;
            for (;;) {
                int l = 2 * i + 1, r = 2 * i + 2, low = (l < hn && Ls[heap[l]] < Ls[heap[i]]) ? l : i;
                if (r < hn && Ls[heap[r]] < Ls[heap[low]]) low = r;
                if (low == i) break;
                int tmp = heap[low]; heap[low] = heap[i]; heap[i] = tmp; i = low;
            }
        }
        if (min_len < 0) min_len = Ls[top];
        if (Ls[top] > 1.6 * min_len && idx > 2) break;
        int T = sample_word(&ws[top], tb->maxc, 0.1 * tb->maxc, pose, buf);
        out->ntried += 1; out->T_last = T;
        if (traj_valid(buf, T, bounds, so)) {
            out->found = 1; out->nseg = ws[top].n; out->L = Ls[top];
            memcpy(out->types, ws[top].types, 5);
            for (int i = 0; i < 5; ++i) out->len[i] = i < ws[top].n ? ws[top].len[i] / tb->maxc : 0.0;
            for (int i = ws[top].n; i < 5; ++i) out->types[i] = 255;
            return;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    /* scene (read-only), all [n][...] */
    const double *start, *dest, *bounds, *obs;
    const int *nverts;
    /* state (in/out) */
    double *pose, *accum;
    int *t;
    /* outputs */
    double *lidar, *mask, *target, *reward, *reward_info;
    int *mask_steps, *status, *substeps, *retreated;
    int *rs_found, *rs_nseg, *rs_ncand, *rs_ntried, *rs_T_last, *rs_err;
    uint8_t *rs_types;
    double *rs_len, *rs_L;
} orc_io;

/* ENV_COLLIDE (configs.py:79, default False): 1 = a collision on the first substep ends the episode (:264-267, :279-282) */
static int g_env_collide = 0;
void orc_set_env_collide(int on) { g_env_collide = on; }

static double angle_diff(double a1, double a2) { /* car_parking_base.py:203-206 */
    double d = acos(cos(a1 - a2));
    return d < PI / 2 ? d : PI - d;
}

/* One env step for scene i.  action: raw policy output in [-1,1]^2 (NULL = the reset step).
   stages: bit0 = observation (lidar+mask+target), bit1 = Reeds-Shepp search. */
static void step_one(int i, const orc_io *io, const double *action, const orc_tables *tb, int stages, traj_buf *buf) {
    const double *dest = io->dest + 3 * i, *bounds = io->bounds + 4 * i, *start = io->start + 3 * i;
    scene_obs so = {io->obs + (size_t)i * MAX_OBS * MAX_V * 2, io->nverts + i * MAX_OBS};
    double *pose = io->pose + 3 * i;
    double prev[3] = {pose[0], pose[1], pose[2]};
    double bx[4], by[4], dbx[4], dby[4];
    int arrive = 0, collide = 0, nsub = 0, nret = 0;
    make_box(dest, dbx, dby);
    double dest_area = shoelace(dbx, dby, 4);
    if (action) {
        /* env_wrapper.py:37-50 — clip, a*(hi-lo)/2 + (hi+lo)/2 with float32 bounds 0.75 / 2.5 */
        double a0 = fmin(fmax(action[0], -1.0), 1.0), a1 = fmin(fmax(action[1], -1.0), 1.0);
        double steer = a0 * 0.75 + 0.0, speed = a1 * 2.5 + 0.0;
        for (int s = 0; s < 10; ++s) { /* car_parking_base.py:259-271 */
            double keep[3] = {pose[0], pose[1], pose[2]};
            ks_substep(pose, steer, speed);
            ++nsub;
            make_box(pose, bx, by);
            if (clip_area(bx, by, dbx, dby) / dest_area > 0.95) { arrive = 1; break; }
            if (collides(bx, by, so)) {
                pose[0] = keep[0]; pose[1] = keep[1]; pose[2] = keep[2]; ++nret;
                if (s == 0) collide = g_env_collide;
                break;
            }
        }
    }
    io->t[i] += 1;
    int t = io->t[i];
    make_box(pose, bx, by);
    double inter = clip_area(bx, by, dbx, dby);
    int status;
    if (arrive) status = ST_ARRIVED;
    else if (collide) status = ST_COLLIDED; /* :282 */
    else if (collides(bx, by, so)) status = ST_COLLIDED; /* :175-184 */
    else if (pose[0] > bounds[1] || pose[0] < bounds[0] || pose[1] > bounds[3] || pose[1] < bounds[2]) status = ST_OUTBOUND;
    else if (inter / dest_area > 0.95) status = ST_ARRIVED;
    else if (t > 200) status = ST_OUTTIME;
    else status = ST_CONTINUE;

    double ri[5] = {0, 0, 0, 0, 0};
    if (status == ST_CONTINUE) { /* :186-227 */
        ri[0] = -tanh((double)t / (10 * 200));
        double dd = hypot(pose[0] - dest[0], pose[1] - dest[1]), pd = hypot(prev[0] - dest[0], prev[1] - dest[1]);
        double norm = fmax(hypot(dest[0] - start[0], dest[1] - start[1]), 10.0);
        ri[2] = pd / norm - dd / norm;
        ri[3] = angle_diff(prev[2], dest[2]) / PI - angle_diff(pose[2], dest[2]) / PI;
        double u = inter / (2 * dest_area - inter);
        if (u < io->accum[i]) u = 0; else { double p = io->accum[i]; io->accum[i] = u; u -= p; }
        ri[4] = u;
    }
    double reward; /* env_wrapper.py:10-35 */
    if (status == ST_CONTINUE) { reward = 0; reward += 1 * ri[0]; reward += 0 * ri[1]; reward += 5 * ri[2]; reward += 0 * ri[3]; reward += 10 * ri[4]; }
    else if (status == ST_OUTBOUND) reward = -50;
    else if (status == ST_OUTTIME) reward = -1;
    else if (status == ST_ARRIVED) reward = 50;
    else reward = -50;
    reward *= 0.1;
    io->status[i] = status; io->reward[i] = reward;
    memcpy(io->reward_info + 5 * i, ri, sizeof ri);
    io->substeps[i] = nsub; io->retreated[i] = nret;

    if (stages & 1) {
        double *lid = io->lidar + (size_t)N_RAY * i;
        lidar(pose, so, tb, lid);
        action_mask(lid, tb, io->mask + (size_t)N_ACT * i, io->mask_steps + (size_t)N_ACT * i);
        double *tg = io->target + 5 * i; /* :372-381 */
        double ddx = dest[0] - pose[0], ddy = dest[1] - pose[1];
        double rel = atan2(ddy, ddx) - pose[2], relh = dest[2] - pose[2];
        tg[0] = sqrt(pow(ddx, 2.0) + pow(ddy, 2.0)); tg[1] = cos(rel); tg[2] = sin(rel); tg[3] = cos(relh); tg[4] = cos(relh);
    }
    rs_result rr;
    memset(&rr, 0, sizeof rr); memset(rr.types, 255, 5);
    if ((stages & 2) && t > 1 && status == ST_CONTINUE && hypot(pose[0] - dest[0], pose[1] - dest[1]) < 10.0) /* :293-297 */
        find_rs_path(pose, dest, bounds, so, tb, buf, &rr);
    io->rs_found[i] = rr.found; io->rs_nseg[i] = rr.nseg; io->rs_ncand[i] = rr.ncand; io->rs_ntried[i] = rr.ntried;
    io->rs_T_last[i] = rr.T_last; io->rs_err[i] = rr.err; io->rs_L[i] = rr.L;
    memcpy(io->rs_types + 5 * i, rr.types, 5); memcpy(io->rs_len + 5 * i, rr.len, sizeof rr.len);
}

/* Batched entry point.  action == NULL runs the reset step (no motion) for every scene.
   act_mask (optional, [n]): 0 = treat scene i as a reset step (action ignored). */
int orc_step(int n, const orc_io *io, const double *action, const uint8_t *has_action, const orc_tables *tb,
             int stages, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        traj_buf buf = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < n; ++i) {
            const double *a = action && (!has_action || has_action[i]) ? action + 2 * i : 0;
            step_one(i, io, a, tb, stages, &buf);
        }
        free(buf.x); free(buf.y); free(buf.yaw);
    }
    return 0;
}

/* Unit entry points for the tests ---------------------------------------------------------- */
int orc_rs_all_paths(const double q0[3], const double q1[3], double maxc, int *nseg, uint8_t *types, double *lens,
                     double *Ls, int *T, double *csum, double *head, double *tail) {
    rs_word ws[MAX_PATHS];
    traj_buf buf = {0, 0, 0, 0};
    int err = 0, nw = enumerate_words(q0, q1, maxc, ws, &err);
    for (int k = 0; k < nw; ++k) {
        nseg[k] = ws[k].n; Ls[k] = ws[k].L / maxc;
        for (int i = 0; i < 5; ++i) { types[5 * k + i] = i < ws[k].n ? ws[k].types[i] : 255; lens[5 * k + i] = i < ws[k].n ? ws[k].len[i] / maxc : 0.0; }
        int n = sample_word(&ws[k], maxc, 0.1 * maxc, q0, &buf);
        T[k] = n;
        double sx = 0, sy = 0, sw = 0;
        for (int i = 0; i < n; ++i) { sx += buf.x[i]; sy += buf.y[i]; sw += buf.yaw[i]; }
        csum[3 * k] = sx; csum[3 * k + 1] = sy; csum[3 * k + 2] = sw;
        for (int j = 0; j < 3 && j < n; ++j) {
            head[9 * k + 3 * j] = buf.x[j]; head[9 * k + 3 * j + 1] = buf.y[j]; head[9 * k + 3 * j + 2] = buf.yaw[j];
            tail[9 * k + 3 * j] = buf.x[n - 1 - j]; tail[9 * k + 3 * j + 1] = buf.y[n - 1 - j]; tail[9 * k + 3 * j + 2] = buf.yaw[n - 1 - j];
        }
    }
    free(buf.x); free(buf.y); free(buf.yaw);
    return err ? -nw - 1 : nw;
}
int orc_orient(double ax, double ay, double bx, double by, double cx, double cy) { return orient(ax, ay, bx, by, cx, cy); }
int orc_seg_hit(const double *p) { return seg_hit(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]); }
double orc_clip_area(const double *sx, const double *sy, const double *cx, const double *cy) { return clip_area(sx, sy, cx, cy); }
int orc_sizeof_io(void) { return (int)sizeof(orc_io); }
int orc_max_obs(void) { return MAX_OBS; }
