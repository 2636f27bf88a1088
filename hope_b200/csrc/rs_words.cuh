// rs_words.cuh — Reeds-Shepp word enumeration (reeds_shepp.py:35-76, 79-449, 540-557): the 46 candidate words of a
// start -> goal query in the reference's order, with its admission rule.  Plain functions over float64 (no thread or
// memory-space dependence) so that tests/rs_host_harness.cpp can compile the very same code with g++ and replay the
// reference's known answers through it; included by hope_kernels.cu inside namespace hope.
#pragma once

struct RsWord {  // one admitted word, lengths in curvature-normalised units
    double len[HOPE_RS_MAX_SEG];
    double L;        // sum |len|, normalised
    uint8_t types[HOPE_RS_MAX_SEG];
    uint8_t n;
    uint8_t pad[2];
};

#ifndef HOPE_CONSTANT
#define HOPE_CONSTANT __constant__
#endif

struct Tuv { double t, u, v; };

__device__ bool w_SLS(double x, double y, double phi, Tuv &o) {  // reeds_shepp.py:133-149
    phi = rs_M(phi);
    if (y != 0.0 && 0.0 < phi && phi < HOPE_PI * 0.99 && (y > 0.0 || y < 0.0)) {
        double tp = tan(phi), th = tan(phi / 2.0);
        double xd = -y / tp + x;
        double dxx = x - xd;
        double r = sqrt(dxx * dxx + y * y);
        o.t = xd - th; o.u = phi; o.v = (y > 0.0 ? r : -r) - th;
        return true;
    }
    return false;
}
// One "frame" = one reflection (x, y, phi) -> (+-x, +-y, +-phi) of the normalised goal.  Every word formula
// starts from one of two points, A = (x - sin phi, y - 1 + cos phi) or B = (x + sin phi, y - 1 - cos phi), in polar
// form; the reference recomputes sin/cos/hypot/atan2 inside each of its 44 calls, here each frame does it once.
// sin(-phi) = -sin(phi) and cos(-phi) = cos(phi) hold exactly for the libdevice routines, and hypot is symmetric
// in its arguments and their signs, so the shared values are the ones each formula would have computed.
struct Frame {
    double x, y, phi;
    double xi, eta;         // B
    double rA, tA;          // |A|, atan2(A.y, A.x)
    double rB, tB, tB2;     // |B|, atan2(B.y, B.x), atan2(B.x, -B.y)
};
__device__ __forceinline__ void make_frame(Frame &F, double x, double y, double phi, double s, double c, bool forward) {
    F.x = x; F.y = y; F.phi = phi;
    const double ax = x - s, ay = y - 1.0 + c;
    F.xi = x + s; F.eta = y - 1.0 - c;
    F.rA = hypot(ax, ay); F.tA = atan2(ay, ax);
    F.rB = hypot(F.xi, F.eta);
    F.tB2 = atan2(F.xi, -F.eta);
    F.tB = forward ? atan2(F.eta, F.xi) : 0.0;  // only LSR needs it, and LSR has no "backwards" variant
}
__device__ bool w_LSL(const Frame &F, Tuv &o) {  // :79-87
    const double t = F.tA;
    if (t >= 0.0) {
        double v = rs_M(F.phi - t);
        if (v >= 0.0) { o.t = t; o.u = F.rA; o.v = v; return true; }
    }
    return false;
}
__device__ bool w_LSR(const Frame &F, Tuv &o) {  // :90-103
    double u1 = F.rB;
    const double t1 = F.tB;
    u1 = u1 * u1;
    if (u1 >= 4.0) {
        double u = sqrt(u1 - 4.0), th = atan2(2.0, u), t = rs_M(t1 + th), v = rs_M(t - F.phi);
        if (t >= 0.0 && v >= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
    }
    return false;
}
__device__ bool w_LRL(const Frame &F, Tuv &o) {  // :106-117
    const double u1 = F.rA, t1 = F.tA;
    if (u1 <= 4.0) {
        double u = -2.0 * asin(0.25 * u1), t = rs_M(t1 + 0.5 * u + HOPE_PI), v = rs_M(F.phi - t + u);
        if (t >= 0.0 && u <= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
    }
    return false;
}
__device__ void tau_omega(double u, double v, double xi, double eta, double phi, double &tau, double &omega) {  // :228-243
    double delta = rs_M(u - v);
    double su, cu, sd, cd;
    sincos(u, &su, &cu);
    sincos(delta, &sd, &cd);
    double A = su - sd, B = cu - cd - 1.0;
    double t1 = atan2(eta * A - xi * B, xi * A + eta * B);
    double t2 = 2.0 * (cd - cos(v) - cu) + 3.0;
    tau = t2 < 0 ? rs_M(t1 + HOPE_PI) : rs_M(t1);
    omega = rs_M(tau - u + v - phi);
}
__device__ bool w_LRLRn(const Frame &F, Tuv &o) {  // :246-257
    const double xi = F.xi, eta = F.eta, rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
    if (rho <= 1.0) {
        double u = acos(rho), t, v;
        tau_omega(u, -u, xi, eta, F.phi, t, v);
        if (t >= 0.0 && v <= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
    }
    return false;
}
__device__ bool w_LRLRp(const Frame &F, Tuv &o) {  // :260-272
    const double xi = F.xi, eta = F.eta, rho = (20.0 - xi * xi - eta * eta) / 16.0;
    if (0.0 <= rho && rho <= 1.0) {
        double u = -acos(rho);
        if (u >= -0.5 * HOPE_PI) {
            double t, v;
            tau_omega(u, u, xi, eta, F.phi, t, v);
            if (t >= 0.0 && v >= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
        }
    }
    return false;
}
__device__ bool w_LRSR(const Frame &F, Tuv &o) {  // :311-323  R(-eta, xi)
    const double rho = F.rB, theta = F.tB2;
    if (rho >= 2.0) {
        double t = theta, u = 2.0 - rho, v = rs_M(t + 0.5 * HOPE_PI - F.phi);
        if (t >= 0.0 && u <= 0.0 && v <= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
    }
    return false;
}
__device__ bool w_LRSL(const Frame &F, Tuv &o) {  // :326-339
    const double rho = F.rA, theta = F.tA;
    if (rho >= 2.0) {
        double r = sqrt(rho * rho - 4.0), u = 2.0 - r, t = rs_M(theta + atan2(r, -2.0)), v = rs_M(F.phi - 0.5 * HOPE_PI - t);
        if (t >= 0.0 && u <= 0.0 && v <= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
    }
    return false;
}
__device__ bool w_LRSLR(const Frame &F, Tuv &o) {  // :414-429
    const double xi = F.xi, eta = F.eta, rho = F.rB;
    if (rho >= 2.0) {
        double u = 4.0 - sqrt(rho * rho - 4.0);
        if (u <= 0.0) {
            double t = rs_M(atan2((4.0 - u) * xi - 2.0 * eta, -2.0 * xi + (u - 4.0) * eta)), v = rs_M(t - F.phi);
            if (t >= 0.0 && v >= 0.0) { o.t = t; o.u = u; o.v = v; return true; }
        }
    }
    return false;
}

// Families in generate_path order (:549-555).  `lay` says where (t,u,v) land in the length vector.
enum { LAY_TUV, LAY_VUT, LAY_T_U_mU_V, LAY_T_U_U_V, LAY_T_H_U_V, LAY_V_U_H_T, LAY_T_H_U_H_V };
enum { F_LSL, F_LSR, F_LRL, F_LRLRn, F_LRLRp, F_LRSL, F_LRSR, F_LRSLR };
struct Family { uint8_t fn, back, lay, n, ty[5]; };
HOPE_CONSTANT Family c_families[11] = {
    {F_LSL, 0, LAY_TUV, 3, {1, 0, 1, 255, 255}},        {F_LSR, 0, LAY_TUV, 3, {1, 0, 2, 255, 255}},
    {F_LRL, 0, LAY_TUV, 3, {1, 2, 1, 255, 255}},        {F_LRL, 1, LAY_VUT, 3, {1, 2, 1, 255, 255}},
    {F_LRLRn, 0, LAY_T_U_mU_V, 4, {1, 2, 1, 2, 255}},   {F_LRLRp, 0, LAY_T_U_U_V, 4, {1, 2, 1, 2, 255}},
    {F_LRSL, 0, LAY_T_H_U_V, 4, {1, 2, 0, 1, 255}},     {F_LRSR, 0, LAY_T_H_U_V, 4, {1, 2, 0, 2, 255}},
    {F_LRSL, 1, LAY_V_U_H_T, 4, {1, 0, 2, 1, 255}},     {F_LRSR, 1, LAY_V_U_H_T, 4, {2, 0, 2, 1, 255}},
    {F_LRSLR, 0, LAY_T_H_U_H_V, 5, {1, 2, 0, 1, 2}},
};

struct WordList {
    double len[MAXW][HOPE_RS_MAX_SEG];
    double L[MAXW];
    uint32_t ty[MAXW];  // 4 bits per segment, 0xF = unused
    uint8_t n[MAXW];
    int count;
};
// set_path (reeds_shepp.py:57-76).  Returns false if the capacity MAXW was hit.
__device__ bool admit(WordList &w, int n, uint32_t ty, const double *len, unsigned long long *counters) {
    for (int k = 0; k < w.count; ++k) {
        if (w.ty[k] != ty) continue;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = s + (w.len[k][i] - len[i]);
        if (s <= 0.01) return true;  // near-duplicate of an earlier word of the same type
    }
    double L = 0.0;
    for (int i = 0; i < n; ++i) L = L + fabs(len[i]);
    if (L >= 1000.0) return true;
    if (!(L >= 0.001)) { if (counters) atomicAdd(counters + 4, 1ull); return true; }  // the reference asserts here (:73)
    if (w.count == MAXW) { if (counters) atomicAdd(counters + 3, 1ull); return false; }
    int k = w.count++;
    for (int i = 0; i < HOPE_RS_MAX_SEG; ++i) w.len[k][i] = i < n ? len[i] : 0.0;
    w.L[k] = L; w.ty[k] = ty; w.n[k] = (uint8_t)n;
    return true;
}

// generate_path + calc_all_paths for the query (sx, sy, sh) -> (gx, gy, gh) at curvature maxc: fills `w` with the
// admitted words in the reference's insertion order (lengths in curvature-normalised units).  Returns false if the
// capacity MAXW was hit.
__device__ bool enumerate_words(double sx, double sy, double sh, double gx, double gy, double gh, double maxc, WordList &w,
                                unsigned long long *counters) {
    // generate_path :540-547
    double dx = gx - sx, dy = gy - sy, phi = gh - sh;
    double c, s;
    sincos(sh, &s, &c);
    const double x = (c * dx + s * dy) * maxc, y = (-s * dx + c * dy) * maxc;

    w.count = 0;
    Tuv o;
    double len[5];
    bool room = true;
    // SCS :120-130
    if (w_SLS(x, y, phi, o)) { len[0] = o.t; len[1] = o.u; len[2] = o.v; room &= admit(w, 3, 0xFF000u | 0x010u, len, counters); }
    if (w_SLS(x, -y, -phi, o)) { len[0] = o.t; len[1] = o.u; len[2] = o.v; room &= admit(w, 3, 0xFF000u | 0x020u, len, counters); }
    double sp, cp;
    sincos(phi, &sp, &cp);
    const double xb = x * cp + y * sp, yb = x * sp - y * cp;  // :206-207, :376-377
    // the four reflections (x,y,phi) (-x,y,-phi) (x,-y,-phi) (-x,-y,phi) of the goal and of the "backwards" goal
    Frame frames[8];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const double X = b ? xb : x, Y = b ? yb : y;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool neg = (r == 1 || r == 2);
            make_frame(frames[4 * b + r], (r & 1) ? -X : X, (r & 2) ? -Y : Y, neg ? -phi : phi, neg ? -sp : sp, cp, b == 0);
        }
    }
    for (int f = 0; f < 11; ++f) {
        const Family F = c_families[f];
        for (int r = 0; r < 4; ++r) {
            const Frame &fr = frames[4 * F.back + r];
            bool ok;
            switch (F.fn) {
            case F_LSL: ok = w_LSL(fr, o); break;
            case F_LSR: ok = w_LSR(fr, o); break;
            case F_LRL: ok = w_LRL(fr, o); break;
            case F_LRLRn: ok = w_LRLRn(fr, o); break;
            case F_LRLRp: ok = w_LRLRp(fr, o); break;
            case F_LRSL: ok = w_LRSL(fr, o); break;
            case F_LRSR: ok = w_LRSR(fr, o); break;
            default: ok = w_LRSLR(fr, o); break;
            }
            if (!ok) continue;
            const double H = -0.5 * HOPE_PI;
            len[3] = len[4] = 0.0;
            switch (F.lay) {
            case LAY_TUV: len[0] = o.t; len[1] = o.u; len[2] = o.v; break;
            case LAY_VUT: len[0] = o.v; len[1] = o.u; len[2] = o.t; break;
            case LAY_T_U_mU_V: len[0] = o.t; len[1] = o.u; len[2] = -o.u; len[3] = o.v; break;
            case LAY_T_U_U_V: len[0] = o.t; len[1] = o.u; len[2] = o.u; len[3] = o.v; break;
            case LAY_T_H_U_V: len[0] = o.t; len[1] = H; len[2] = o.u; len[3] = o.v; break;
            case LAY_V_U_H_T: len[0] = o.v; len[1] = o.u; len[2] = H; len[3] = o.t; break;
            default: len[0] = o.t; len[1] = H; len[2] = o.u; len[3] = H; len[4] = o.v; break;
            }
            if (r & 1) for (int k = 0; k < F.n; ++k) len[k] = -len[k];
            uint32_t ty = 0;
            for (int k = 0; k < 5; ++k) {
                uint32_t b = F.ty[k] == 255 ? 0xFu : F.ty[k];
                if ((r & 2) && (b == 1 || b == 2)) b = 3 - b;  // reflected words swap L and R
                ty |= b << (4 * k);
            }
            room &= admit(w, F.n, ty, len, counters);
        }
    }
    return room;
}
