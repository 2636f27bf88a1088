// scene_gen.cu — procedural parking scenes (bay / parallel): one generator, compiled for the host
// (multi-threaded, hope_generate_scenes) and for the device (one thread per scene, k_generate_scenes,
// scope row f4: auto-reset never leaves the GPU).
//
// Same construction and the same distributions as the reference generator
// (src/env/parking_map_normal.py:25-494), with an own counter-based RNG stream
// per scene (seed + index) instead of numpy's global generator: bit-identical regeneration of
// the reference's scenes is not required (only step parity on identical scenes), and an
// independent stream per scene makes generation order- and thread-count-invariant.
// Rejected attempts are retried in a loop (the reference recurses, :242-246, :454-457).
// Host and device draw the same integers; their libm/libdevice transcendentals can differ in the last
// ulp, so a device scene is not bit-identical to the host scene of the same seed.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/hope_b200.h"
#include "scene_gen.h"

#define HD __host__ __device__

namespace hope_scene {

HD inline double dmin(double a, double b) { return a < b ? a : b; }
HD inline double dmax(double a, double b) { return a > b ? a : b; }

constexpr double PI = 3.141592653589793;
// configs.py:13-70
constexpr double WHEEL_BASE = 2.8, FRONT_HANG = 0.96, REAR_HANG = 0.93, WIDTH = 1.94;
constexpr double LENGTH = WHEEL_BASE + FRONT_HANG + REAR_HANG;
constexpr double MIN_DIST_TO_OBST = 0.1;
constexpr double P_HUGE = 0.5, P_KEEP_EXTRA = 0.7;  // parking_map_normal.py:21-23
// level: 0 Normal, 1 Complex, 2 Extrem
HD inline double min_lot_len(int l) { return l == 0 ? LENGTH * 1.25 : (l == 1 ? LENGTH + 0.9 : LENGTH + 0.6); }
HD inline double max_lot_len(int l) { return l == 0 ? LENGTH * 1.25 + 0.5 : (l == 1 ? LENGTH * 1.25 : LENGTH + 0.9); }
HD inline double min_lot_width(int l) { return l == 0 ? WIDTH + 0.85 : WIDTH + 0.4; }
HD inline double max_lot_width(int l) { return l == 0 ? WIDTH + 1.2 : WIDTH + 0.85; }
HD inline double para_wall_dist(int l) { return l == 0 ? 4.5 : (l == 1 ? 4.0 : 3.5); }
HD inline double bay_wall_dist(int l) { return l == 0 ? 7.0 : 6.0; }
HD inline int n_obstacle(int l) { return l == 0 ? 3 : (l == 1 ? 5 : 8); }

struct Rng {  // xoshiro256** seeded by splitmix64
    uint64_t s[4];
    bool has_spare = false;
    double spare = 0;
    HD static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    HD explicit Rng(uint64_t seed) { for (int k = 0; k < 4; ++k) s[k] = splitmix(seed); }
    HD static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    HD uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    HD double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    HD double normal() {  // Marsaglia polar
        if (has_spare) { has_spare = false; return spare; }
        double u, v, q;
        do { u = 2 * uniform() - 1; v = 2 * uniform() - 1; q = u * u + v * v; } while (q >= 1.0 || q == 0.0);
        double f = sqrt(-2.0 * log(q) / q);
        spare = v * f; has_spare = true;
        return u * f;
    }
    HD double gauss(double mean, double sd, double lo, double hi) { return dmin(dmax(normal() * sd + mean, lo), hi); }  // :25-27
    HD double between(double lo, double hi) { return uniform() * (hi - lo) + lo; }                                              // :29-31
};

struct P2 { double x, y; };
struct Ring { P2 v[4]; int n = 4; };

HD Ring car_box(double x, double y, double yaw) {  // vehicle.py:32-36 corner order rb, rf, lf, lb
    const double bx[4] = {-REAR_HANG, FRONT_HANG + WHEEL_BASE, FRONT_HANG + WHEEL_BASE, -REAR_HANG};
    const double by[4] = {-WIDTH / 2, -WIDTH / 2, WIDTH / 2, WIDTH / 2};
    double c = cos(yaw), s = sin(yaw);
    Ring r;
    for (int i = 0; i < 4; ++i) r.v[i] = {c * bx[i] + (-s) * by[i] + x, s * bx[i] + c * by[i] + y};
    return r;
}

// exact orientation sign (static filter, then expansion arithmetic)
HD void two_sum(double a, double b, double &s, double &e) { s = a + b; double bv = s - a, av = s - bv; e = (a - av) + (b - bv); }
HD void two_prod(double a, double b, double &p, double &e) { p = a * b; e = fma(a, b, -p); }
HD int orient(P2 a, P2 b, P2 c) {
    const double eps = 1.1102230246251565e-16;
    double l = (a.x - c.x) * (b.y - c.y), r = (a.y - c.y) * (b.x - c.x), det = l - r;
    double bound = (3.0 + 16.0 * eps) * eps * (fabs(l) + fabs(r));
    if (det > bound) return 1;
    if (det < -bound) return -1;
    double t[12], e[12];
    two_prod(a.x, b.y, t[0], t[1]);  two_prod(-a.x, c.y, t[2], t[3]);  two_prod(-c.x, b.y, t[4], t[5]);
    two_prod(-a.y, b.x, t[6], t[7]); two_prod(a.y, c.x, t[8], t[9]);   two_prod(c.y, b.x, t[10], t[11]);
    int n = 0;
    for (int k = 0; k < 12; ++k) {
        double q = t[k];
        for (int i = 0; i < n; ++i) { double s, rr; two_sum(q, e[i], s, rr); e[i] = rr; q = s; }
        e[n++] = q;
    }
    for (int i = n - 1; i >= 0; --i) if (e[i] != 0.0) return e[i] > 0 ? 1 : -1;
    return 0;
}
HD bool in_span(P2 p, P2 a, P2 b) {
    return dmin(a.x, b.x) <= p.x && p.x <= dmax(a.x, b.x) && dmin(a.y, b.y) <= p.y && p.y <= dmax(a.y, b.y);
}
HD bool seg_touch(P2 p1, P2 p2, P2 q1, P2 q2) {
    if (dmax(p1.x, p2.x) < dmin(q1.x, q2.x) || dmax(q1.x, q2.x) < dmin(p1.x, p2.x)) return false;
    if (dmax(p1.y, p2.y) < dmin(q1.y, q2.y) || dmax(q1.y, q2.y) < dmin(p1.y, p2.y)) return false;
    int o1 = orient(p1, p2, q1), o2 = orient(p1, p2, q2), o3 = orient(q1, q2, p1), o4 = orient(q1, q2, p2);
    if (o1 * o2 < 0 && o3 * o4 < 0) return true;
    return (o1 == 0 && in_span(q1, p1, p2)) || (o2 == 0 && in_span(q2, p1, p2)) || (o3 == 0 && in_span(p1, q1, q2)) ||
           (o4 == 0 && in_span(p2, q1, q2));
}
HD bool rings_touch(const Ring &a, const Ring &b) {  // LinearRing.intersects(LinearRing): boundaries only
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < b.n; ++j)
            if (seg_touch(a.v[i], a.v[(i + 1) % a.n], b.v[j], b.v[(j + 1) % b.n])) return true;
    return false;
}
HD double point_seg(P2 p, P2 a, P2 b) {
    double dx = b.x - a.x, dy = b.y - a.y, l2 = dx * dx + dy * dy;
    if (l2 == 0.0) return hypot(p.x - a.x, p.y - a.y);
    double r = ((p.x - a.x) * dx + (p.y - a.y) * dy) / l2;
    if (r <= 0.0) return hypot(p.x - a.x, p.y - a.y);
    if (r >= 1.0) return hypot(p.x - b.x, p.y - b.y);
    return fabs((a.y - p.y) * dx - (a.x - p.x) * dy) / sqrt(l2);
}
HD double ring_gap(const Ring &a, const Ring &b) {  // ring.distance(ring)
    if (rings_touch(a, b)) return 0.0;
    double best = INFINITY;
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < b.n; ++j) {
            P2 a0 = a.v[i], a1 = a.v[(i + 1) % a.n], b0 = b.v[j], b1 = b.v[(j + 1) % b.n];
            best = dmin(dmin(best, point_seg(a0, b0, b1)), dmin(point_seg(a1, b0, b1), dmin(point_seg(b0, a0, a1), point_seg(b1, a0, a1))));
        }
    return best;
}
HD P2 polar_offset(Rng &g, P2 o, double amin, double amax, double rmin, double rmax) {  // get_rand_pos :33-38
    double ang = g.gauss((amax + amin) / 2, (amax - amin) / 4, amin, amax);
    double rad = g.gauss((rmin + rmax) / 2, (rmax - rmin) / 4, rmin, rmax);
    return {o.x + cos(ang) * rad, o.y + sin(ang) * rad};
}

constexpr int MAX_RINGS = 24;  // back + 2 neighbours + 6 extras + 8 far-side = 17 at most; HOPE_MAX_OBS is checked on output

struct Scene {
    double start[3], dest[3];
    Ring obs[MAX_RINGS];
    int n_obs, case_id;
    HD void push(const Ring &r) { if (n_obs < MAX_RINGS) obs[n_obs] = r; ++n_obs; }
};

struct Parked { Ring box; double y; };

// One attempt at a bay (parallel == false, :40-246) or parallel (:248-457) case.
HD bool attempt(Rng &g, int level, bool parallel, Scene &sc) {
    const double half = parallel ? 18.0 : 15.0;
    const double car_span = parallel ? LENGTH : WIDTH;  // neighbour pitch along the kerb
    const double max_space = parallel ? max_lot_len(level) - LENGTH : max_lot_width(level) - WIDTH;
    const double min_space = parallel ? min_lot_len(level) - LENGTH : min_lot_width(level) - WIDTH;
    const double wall_dist = parallel ? para_wall_dist(level) : bay_wall_dist(level);
    const double yaw_mean = parallel ? 0.0 : PI / 2, yaw_lo = parallel ? -PI / 12 : PI * 5 / 12, yaw_hi = parallel ? PI / 12 : PI * 7 / 12;
    const int n_extra = parallel ? 2 : 3;
    bool ok = true;
    // corners whose y decides how close a parked car may sit to the back wall: (rb, lb) bay, (rb, rf) parallel
    auto kerb_clearance = [&](const Ring &b) { return -dmin(b.v[0].y, parallel ? b.v[1].y : b.v[3].y) + MIN_DIST_TO_OBST; };
    auto parked_car = [&](double x) {
        double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
        double ymin = kerb_clearance(car_box(x, 0.0, yaw));
        double y = g.gauss(ymin + 0.4, 0.2, ymin, ymin + 0.8);
        Parked r{car_box(x, y, yaw), y};
        return r;
    };
    sc.n_obs = 0;
    Ring extra[6];
    int n_kept = 0;
    Ring back;
    back.v[0] = {half, 0}; back.v[1] = {half, -1}; back.v[2] = {-half, -1}; back.v[3] = {-half, 0};
    // destination slot
    double dest_yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
    double dclear = kerb_clearance(car_box(0, 0, dest_yaw));
    double dest_x = 0.0, dest_y = g.gauss(dclear + 0.4, 0.2, dclear, dclear + 0.8);
    Ring dest_box = car_box(dest_x, dest_y, dest_yaw);
    const P2 rb = dest_box.v[0], rf = dest_box.v[1], lf = dest_box.v[2], lb = dest_box.v[3];
    // left neighbour
    Ring left;
    {
        double dhi = max_space / 5 * 4, dlo = (parallel ? min_space : max_space) / 5 * 1;
        if (g.uniform() < P_HUGE) {
            P2 a = polar_offset(g, parallel ? lb : lf, PI * 11 / 12, PI * 13 / 12, dlo, dhi);
            P2 b = polar_offset(g, parallel ? rb : lb, PI * 11 / 12, PI * 13 / 12, dlo, dhi);
            left.v[0] = a; left.v[1] = b; left.v[2] = {-half, 0}; left.v[3] = {-half, a.y};
        } else {
            double cx = 0.0 - (car_span + g.between(dlo, dhi));
            Parked c = parked_car(cx);
            left = c.box;
            double cy = c.y;
            for (int k = 0; k < n_extra; ++k) {
                cx -= (car_span + MIN_DIST_TO_OBST + g.between(dlo, dhi));
                cy += g.gauss(0, 0.05, -0.1, 0.1);
                double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
                Ring e = car_box(cx, cy, yaw);
                if (g.uniform() < P_KEEP_EXTRA) extra[n_kept++] = e;
            }
        }
    }
    // right neighbour: the gap budget left over after the left one
    double gap_l = ring_gap(dest_box, left);
    Ring right;
    {
        double dlo = dmax(min_space - gap_l, 0.0) + MIN_DIST_TO_OBST, dhi = dmax(max_space - gap_l, 0.0) + MIN_DIST_TO_OBST;
        if (g.uniform() < P_HUGE) {
            P2 a = polar_offset(g, parallel ? lf : rf, -PI / 12, PI / 12, dlo, dhi);
            P2 b = polar_offset(g, parallel ? rf : rb, -PI / 12, PI / 12, dlo, dhi);
            right.v[0] = {half, a.y}; right.v[1] = {half, 0}; right.v[2] = b; right.v[3] = a;
        } else {
            double cx = 0.0 + (car_span + g.between(dlo, dhi));
            Parked c = parked_car(cx);
            right = c.box;
            double cy = c.y;
            for (int k = 0; k < n_extra; ++k) {
                cx += (car_span + MIN_DIST_TO_OBST + g.between(dlo, dhi));
                cy += g.gauss(0, 0.05, -0.1, 0.1);
                double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
                Ring e = car_box(cx, cy, yaw);
                if (g.uniform() < P_KEEP_EXTRA) extra[n_kept++] = e;
            }
        }
    }
    double gap_r = ring_gap(dest_box, right);
    if (gap_r + gap_l < min_space || gap_r + gap_l > max_space || gap_l < MIN_DIST_TO_OBST || gap_r < MIN_DIST_TO_OBST) ok = false;
    sc.push(back); sc.push(left); sc.push(right);
    for (int k = 0; k < n_kept; ++k) sc.push(extra[k]);
    for (int k = 0; k < sc.n_obs; ++k) if (rings_touch(sc.obs[k], dest_box)) ok = false;
    // far side of the aisle
    double top = -INFINITY;
    for (int k = 0; k < sc.n_obs; ++k) for (int i = 0; i < sc.obs[k].n; ++i) top = dmax(top, sc.obs[k].v[i].y);
    top += MIN_DIST_TO_OBST;
    const int far0 = sc.n_obs;
    if (g.uniform() < 0.2) {
        double y0 = wall_dist + top + MIN_DIST_TO_OBST;
        Ring w;
        w.v[0] = {-half, y0}; w.v[1] = {half, y0}; w.v[2] = {half, y0 + 0.1}; w.v[3] = {-half, y0 + 0.1};
        sc.push(w);
    } else {
        Ring zone;
        zone.v[0] = {-half, wall_dist + top}; zone.v[1] = {half, wall_dist + top}; zone.v[2] = {half, wall_dist + top + 8}; zone.v[3] = {-half, wall_dist + top + 8};
        for (int k = 0; k < n_obstacle(level); ++k) {
            double ox = g.between(-half + 2, half - 2), oy = g.between(wall_dist + top + 2, wall_dist + top + 6);
            Ring o = car_box(ox, oy, g.uniform() * PI * 2);
            for (int i = 0; i < 4; ++i) { o.v[i].x += 0.5 * g.uniform(); o.v[i].y += 0.5 * g.uniform(); }
            if (rings_touch(o, zone)) continue;
            bool clash = false;
            for (int f = far0; f < sc.n_obs; ++f) if (rings_touch(o, sc.obs[f])) { clash = true; break; }
            if (!clash) sc.push(o);
        }
    }
    if (sc.n_obs > MAX_RINGS) return false;
    // start pose: anywhere in the aisle that touches nothing
    double sx, sy, syaw;
    for (int guard = 0;; ++guard) {
        sx = g.between(-half / 2, half / 2);
        sy = g.between(top + 1, wall_dist + top - 1);
        syaw = g.gauss(0, PI / 6, -PI / 2, PI / 2);
        if (g.uniform() < 0.5) syaw += PI;
        Ring sb = car_box(sx, sy, syaw);
        bool free_ = !rings_touch(dest_box, sb);
        for (int k = 0; k < sc.n_obs; ++k) if (rings_touch(sc.obs[k], sb)) free_ = false;
        if (free_) break;
        if (guard > 10000) return false;
    }
    if (parallel && cos(syaw) < 0) {  // :437-442 face the slot the way the car arrives
        double cx = (rb.x + rf.x + lf.x + lb.x) / 4, cy = (rb.y + rf.y + lf.y + lb.y) / 4;
        dest_x = 2 * cx - dest_x; dest_y = 2 * cy - dest_y; dest_yaw += PI;
    }
    if (!ok) return false;
    sc.start[0] = sx; sc.start[1] = sy; sc.start[2] = syaw;
    sc.dest[0] = dest_x; sc.dest[1] = dest_y; sc.dest[2] = dest_yaw;
    sc.case_id = parallel ? 1 : 0;
    return true;
}

// Scene `index` of the stream `seed`: ParkingMapNormal.reset (:474-494) with case_id None.
HD int make_scene(uint64_t seed, uint64_t index, int level, Scene &sc) {
    Rng g(seed + index * 0x9e3779b97f4a7c15ull + 1);
    const bool bay = (g.uniform() > 0.5) && level != 2;
    int tries = 0;
    while (!attempt(g, level, !bay, sc)) if (++tries > 100000) return HOPE_ERR_INVALID;
    if (sc.n_obs > HOPE_MAX_OBS) return HOPE_ERR_CAPACITY;
    return HOPE_OK;
}

HD void scene_bounds(const Scene &sc, double b[4]) {  // :486-489
    b[0] = floor(dmin(sc.start[0], sc.dest[0]) - 10); b[1] = ceil(dmax(sc.start[0], sc.dest[0]) + 10);
    b[2] = floor(dmin(sc.start[1], sc.dest[1]) - 10); b[3] = ceil(dmax(sc.start[1], sc.dest[1]) + 10);
}

// Device generator: one thread per scene, writing the pool arrays of hope_kernels.cu directly
// (obs[P][16][4][2], nv[P][16], aabb[P][16][4], meta[P][24], nobs[P]; layout in DESIGN.md §3).
__global__ void __launch_bounds__(64) k_generate_scenes(int n, int first, int level_or_mix, uint64_t seed, const int *__restrict__ slots,
                                                        const unsigned *__restrict__ episode, hope_params par, double *obs, uint8_t *nv,
                                                        double *aabb, double *meta, int *nobs, int *status) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int slot = slots ? slots[t] : first + t;
    if (slot < 0) return;
    const int level = level_or_mix >= 0 ? level_or_mix : slot % 3;
    const uint64_t index = (uint64_t)slot + (episode ? (uint64_t)episode[slot] << 32 : 0ull);
    Scene sc;
    const int rc = make_scene(seed, index, level, sc);
    if (rc != HOPE_OK) { atomicMin(status, rc); return; }
    double *m = meta + (size_t)slot * 24;
    for (int k = 0; k < 3; ++k) { m[k] = sc.start[k]; m[3 + k] = sc.dest[k]; }
    scene_bounds(sc, m + 6);
    const double c = cos(sc.dest[2]), s = sin(sc.dest[2]), ms = -s;
    double bx[4], by[4], sa = 0.0;
    for (int k = 0; k < 4; ++k) {
        bx[k] = c * par.box_x[k] + ms * par.box_y[k] + sc.dest[0];
        by[k] = s * par.box_x[k] + c * par.box_y[k] + sc.dest[1];
        m[10 + k] = bx[k]; m[14 + k] = by[k];
    }
    for (int k = 0; k < 4; ++k) { int j = (k + 1) & 3; sa += bx[k] * by[j] - bx[j] * by[k]; }
    m[18] = fabs(sa) * 0.5;
    m[19] = dmax(hypot(sc.dest[0] - sc.start[0], sc.dest[1] - sc.start[1]), 10.0);
    m[20] = dmin(dmin(bx[0], bx[1]), dmin(bx[2], bx[3])); m[21] = dmax(dmax(bx[0], bx[1]), dmax(bx[2], bx[3]));
    m[22] = dmin(dmin(by[0], by[1]), dmin(by[2], by[3])); m[23] = dmax(dmax(by[0], by[1]), dmax(by[2], by[3]));
    double *o = obs + (size_t)slot * HOPE_MAX_OBS * HOPE_MAX_VERTS * 2;
    double *bb = aabb + (size_t)slot * HOPE_MAX_OBS * 4;
    uint8_t *v = nv + (size_t)slot * HOPE_MAX_OBS;
    for (int k = 0; k < HOPE_MAX_OBS; ++k) {
        const bool used = k < sc.n_obs;
        v[k] = used ? 4 : 0;
        double xmn = 0, xmx = 0, ymn = 0, ymx = 0;
        for (int j = 0; j < 4; ++j) {
            const double px = used ? sc.obs[k].v[j].x : 0.0, py = used ? sc.obs[k].v[j].y : 0.0;
            o[(k * 4 + j) * 2] = px; o[(k * 4 + j) * 2 + 1] = py;
            if (j == 0) { xmn = xmx = px; ymn = ymx = py; }
            else { xmn = dmin(xmn, px); xmx = dmax(xmx, px); ymn = dmin(ymn, py); ymx = dmax(ymx, py); }
        }
        bb[4 * k] = xmn; bb[4 * k + 1] = xmx; bb[4 * k + 2] = ymn; bb[4 * k + 3] = ymx;
    }
    nobs[slot] = sc.n_obs;
}

}  // namespace hope_scene

namespace hope_scene {
int launch_generate(int n, int first, int level_or_mix, uint64_t seed, const int *d_slots, const unsigned *d_episode, const hope_params &par,
                    double *obs, uint8_t *nv, double *aabb, double *meta, int *nobs, int *d_status, void *stream) {
    k_generate_scenes<<<(n + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(n, first, level_or_mix, seed, d_slots, d_episode, par, obs, nv,
                                                                                   aabb, meta, nobs, d_status);
    return cudaGetLastError() == cudaSuccess ? HOPE_OK : HOPE_ERR_CUDA;
}
}  // namespace hope_scene

extern "C" int hope_generate_scenes(int n, int level, uint64_t seed, int nthreads, double *h_start, double *h_dest, double *h_bounds,
                                    double *h_obs_xy, int32_t *h_nverts, int32_t *h_case_id) {
    using namespace hope_scene;
    if (n <= 0 || level < 0 || level > 2 || !h_start || !h_dest || !h_bounds || !h_obs_xy || !h_nverts) return HOPE_ERR_INVALID;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min(nthreads, n);
    std::vector<int> rc(nthreads, HOPE_OK);
    auto work = [&](int tid) {
        for (int i = tid; i < n; i += nthreads) {
            Scene sc;
            int r = make_scene(seed, (uint64_t)i, level, sc);
            if (r != HOPE_OK) { rc[tid] = r; return; }
            memcpy(h_start + 3 * i, sc.start, 24); memcpy(h_dest + 3 * i, sc.dest, 24);
            scene_bounds(sc, h_bounds + 4 * i);
            double *o = h_obs_xy + (size_t)i * HOPE_MAX_OBS * HOPE_MAX_VERTS * 2;
            int32_t *nv = h_nverts + (size_t)i * HOPE_MAX_OBS;
            memset(o, 0, sizeof(double) * HOPE_MAX_OBS * HOPE_MAX_VERTS * 2);
            memset(nv, 0, sizeof(int32_t) * HOPE_MAX_OBS);
            for (int k = 0; k < sc.n_obs; ++k) {
                nv[k] = sc.obs[k].n;
                for (int j = 0; j < sc.obs[k].n; ++j) { o[(k * HOPE_MAX_VERTS + j) * 2] = sc.obs[k].v[j].x; o[(k * HOPE_MAX_VERTS + j) * 2 + 1] = sc.obs[k].v[j].y; }
            }
            if (h_case_id) h_case_id[i] = sc.case_id;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
    for (auto &t : pool) t.join();
    for (int r : rc) if (r != HOPE_OK) return r;
    return HOPE_OK;
}
