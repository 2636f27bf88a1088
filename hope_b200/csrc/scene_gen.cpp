// scene_gen.cpp — host-side procedural parking scenes (bay / parallel), multi-threaded.
//
// Same construction and the same distributions as the reference generator
// (src/env/parking_map_normal.py:25-494), with an own counter-based RNG stream
// per scene (seed + index) instead of numpy's global generator: bit-identical regeneration of
// the reference's scenes is not required (only step parity on identical scenes), and an
// independent stream per scene makes generation order- and thread-count-invariant.
// Rejected attempts are retried in a loop (the reference recurses, :242-246, :454-457).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/hope_b200.h"

namespace {

const double PI = 3.141592653589793;
// configs.py:13-70
const double WHEEL_BASE = 2.8, FRONT_HANG = 0.96, REAR_HANG = 0.93, WIDTH = 1.94;
const double LENGTH = WHEEL_BASE + FRONT_HANG + REAR_HANG;
const double MIN_DIST_TO_OBST = 0.1;
// index: 0 Normal, 1 Complex, 2 Extrem
const double MIN_LOT_LEN[3] = {LENGTH * 1.25, LENGTH + 0.9, LENGTH + 0.6};
const double MAX_LOT_LEN[3] = {LENGTH * 1.25 + 0.5, LENGTH * 1.25, LENGTH + 0.9};
const double MIN_LOT_WIDTH[3] = {WIDTH + 0.85, WIDTH + 0.4, 0};
const double MAX_LOT_WIDTH[3] = {WIDTH + 1.2, WIDTH + 0.85, 0};
const double PARA_WALL_DIST[3] = {4.5, 4.0, 3.5};
const double BAY_WALL_DIST[3] = {7.0, 6.0, 0};
const int N_OBSTACLE[3] = {3, 5, 8};
const double P_HUGE = 0.5, P_KEEP_EXTRA = 0.7;  // parking_map_normal.py:21-23

struct Rng {  // xoshiro256** seeded by splitmix64
    uint64_t s[4];
    bool has_spare = false;
    double spare = 0;
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double normal() {  // Marsaglia polar
        if (has_spare) { has_spare = false; return spare; }
        double u, v, q;
        do { u = 2 * uniform() - 1; v = 2 * uniform() - 1; q = u * u + v * v; } while (q >= 1.0 || q == 0.0);
        double f = sqrt(-2.0 * log(q) / q);
        spare = v * f; has_spare = true;
        return u * f;
    }
    double gauss(double mean, double sd, double lo, double hi) { return std::min(std::max(normal() * sd + mean, lo), hi); }  // :25-27
    double between(double lo, double hi) { return uniform() * (hi - lo) + lo; }                                              // :29-31
};

struct P2 { double x, y; };
struct Ring { P2 v[4]; int n = 4; };

Ring car_box(double x, double y, double yaw) {  // vehicle.py:32-36 corner order rb, rf, lf, lb
    const double bx[4] = {-REAR_HANG, FRONT_HANG + WHEEL_BASE, FRONT_HANG + WHEEL_BASE, -REAR_HANG};
    const double by[4] = {-WIDTH / 2, -WIDTH / 2, WIDTH / 2, WIDTH / 2};
    double c = cos(yaw), s = sin(yaw);
    Ring r;
    for (int i = 0; i < 4; ++i) r.v[i] = {c * bx[i] + (-s) * by[i] + x, s * bx[i] + c * by[i] + y};
    return r;
}

// exact orientation sign (static filter, then expansion arithmetic)
void two_sum(double a, double b, double &s, double &e) { s = a + b; double bv = s - a, av = s - bv; e = (a - av) + (b - bv); }
void two_prod(double a, double b, double &p, double &e) { p = a * b; e = fma(a, b, -p); }
int orient(P2 a, P2 b, P2 c) {
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
bool in_span(P2 p, P2 a, P2 b) {
    return std::min(a.x, b.x) <= p.x && p.x <= std::max(a.x, b.x) && std::min(a.y, b.y) <= p.y && p.y <= std::max(a.y, b.y);
}
bool seg_touch(P2 p1, P2 p2, P2 q1, P2 q2) {
    if (std::max(p1.x, p2.x) < std::min(q1.x, q2.x) || std::max(q1.x, q2.x) < std::min(p1.x, p2.x)) return false;
    if (std::max(p1.y, p2.y) < std::min(q1.y, q2.y) || std::max(q1.y, q2.y) < std::min(p1.y, p2.y)) return false;
    int o1 = orient(p1, p2, q1), o2 = orient(p1, p2, q2), o3 = orient(q1, q2, p1), o4 = orient(q1, q2, p2);
    if (o1 * o2 < 0 && o3 * o4 < 0) return true;
    return (o1 == 0 && in_span(q1, p1, p2)) || (o2 == 0 && in_span(q2, p1, p2)) || (o3 == 0 && in_span(p1, q1, q2)) ||
           (o4 == 0 && in_span(p2, q1, q2));
}
bool rings_touch(const Ring &a, const Ring &b) {  // LinearRing.intersects(LinearRing): boundaries only
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < b.n; ++j)
            if (seg_touch(a.v[i], a.v[(i + 1) % a.n], b.v[j], b.v[(j + 1) % b.n])) return true;
    return false;
}
double point_seg(P2 p, P2 a, P2 b) {
    double dx = b.x - a.x, dy = b.y - a.y, l2 = dx * dx + dy * dy;
    if (l2 == 0.0) return hypot(p.x - a.x, p.y - a.y);
    double r = ((p.x - a.x) * dx + (p.y - a.y) * dy) / l2;
    if (r <= 0.0) return hypot(p.x - a.x, p.y - a.y);
    if (r >= 1.0) return hypot(p.x - b.x, p.y - b.y);
    return fabs((a.y - p.y) * dx - (a.x - p.x) * dy) / sqrt(l2);
}
double ring_gap(const Ring &a, const Ring &b) {  // ring.distance(ring)
    if (rings_touch(a, b)) return 0.0;
    double best = INFINITY;
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < b.n; ++j) {
            P2 a0 = a.v[i], a1 = a.v[(i + 1) % a.n], b0 = b.v[j], b1 = b.v[(j + 1) % b.n];
            best = std::min(std::min(best, point_seg(a0, b0, b1)), std::min(point_seg(a1, b0, b1), std::min(point_seg(b0, a0, a1), point_seg(b1, a0, a1))));
        }
    return best;
}
P2 polar_offset(Rng &g, P2 o, double amin, double amax, double rmin, double rmax) {  // get_rand_pos :33-38
    double ang = g.gauss((amax + amin) / 2, (amax - amin) / 4, amin, amax);
    double rad = g.gauss((rmin + rmax) / 2, (rmax - rmin) / 4, rmin, rmax);
    return {o.x + cos(ang) * rad, o.y + sin(ang) * rad};
}

struct Scene { double start[3], dest[3]; std::vector<Ring> obs; int case_id; };

// One attempt at a bay (parallel == false, :40-246) or parallel (:248-457) case.
bool attempt(Rng &g, int level, bool parallel, Scene &sc) {
    const double half = parallel ? 18.0 : 15.0;
    const double car_span = parallel ? LENGTH : WIDTH;  // neighbour pitch along the kerb
    const double max_space = parallel ? MAX_LOT_LEN[level] - LENGTH : MAX_LOT_WIDTH[level] - WIDTH;
    const double min_space = parallel ? MIN_LOT_LEN[level] - LENGTH : MIN_LOT_WIDTH[level] - WIDTH;
    const double wall_dist = parallel ? PARA_WALL_DIST[level] : BAY_WALL_DIST[level];
    const double yaw_mean = parallel ? 0.0 : PI / 2, yaw_lo = parallel ? -PI / 12 : PI * 5 / 12, yaw_hi = parallel ? PI / 12 : PI * 7 / 12;
    const int n_extra = parallel ? 2 : 3;
    bool ok = true;
    // corners whose y decides how close a parked car may sit to the back wall: (rb, lb) bay, (rb, rf) parallel
    auto kerb_clearance = [&](const Ring &b) { return -std::min(b.v[0].y, parallel ? b.v[1].y : b.v[3].y) + MIN_DIST_TO_OBST; };
    auto parked_car = [&](double x) {
        double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
        double ymin = kerb_clearance(car_box(x, 0.0, yaw));
        double y = g.gauss(ymin + 0.4, 0.2, ymin, ymin + 0.8);
        struct R { Ring box; double y; } r{car_box(x, y, yaw), y};
        return r;
    };
    std::vector<Ring> obstacles, extra;
    Ring back;
    back.v[0] = {half, 0}; back.v[1] = {half, -1}; back.v[2] = {-half, -1}; back.v[3] = {-half, 0};
    // destination slot
    double dest_yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
    double dmin = kerb_clearance(car_box(0, 0, dest_yaw));
    double dest_x = 0.0, dest_y = g.gauss(dmin + 0.4, 0.2, dmin, dmin + 0.8);
    Ring dest_box = car_box(dest_x, dest_y, dest_yaw);
    const P2 rb = dest_box.v[0], rf = dest_box.v[1], lf = dest_box.v[2], lb = dest_box.v[3];
    // left neighbour
    Ring left;
    {
        double dmax = max_space / 5 * 4, dlo = (parallel ? min_space : max_space) / 5 * 1;
        if (g.uniform() < P_HUGE) {
            P2 a = polar_offset(g, parallel ? lb : lf, PI * 11 / 12, PI * 13 / 12, dlo, dmax);
            P2 b = polar_offset(g, parallel ? rb : lb, PI * 11 / 12, PI * 13 / 12, dlo, dmax);
            left.v[0] = a; left.v[1] = b; left.v[2] = {-half, 0}; left.v[3] = {-half, a.y};
        } else {
            double cx = 0.0 - (car_span + g.between(dlo, dmax));
            auto c = parked_car(cx);
            left = c.box;
            double cy = c.y;
            for (int k = 0; k < n_extra; ++k) {
                cx -= (car_span + MIN_DIST_TO_OBST + g.between(dlo, dmax));
                cy += g.gauss(0, 0.05, -0.1, 0.1);
                double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
                Ring e = car_box(cx, cy, yaw);
                if (g.uniform() < P_KEEP_EXTRA) extra.push_back(e);
            }
        }
    }
    // right neighbour: the gap budget left over after the left one
    double gap_l = ring_gap(dest_box, left);
    Ring right;
    {
        double dlo = std::max(min_space - gap_l, 0.0) + MIN_DIST_TO_OBST, dmax = std::max(max_space - gap_l, 0.0) + MIN_DIST_TO_OBST;
        if (g.uniform() < P_HUGE) {
            P2 a = polar_offset(g, parallel ? lf : rf, -PI / 12, PI / 12, dlo, dmax);
            P2 b = polar_offset(g, parallel ? rf : rb, -PI / 12, PI / 12, dlo, dmax);
            right.v[0] = {half, a.y}; right.v[1] = {half, 0}; right.v[2] = b; right.v[3] = a;
        } else {
            double cx = 0.0 + (car_span + g.between(dlo, dmax));
            auto c = parked_car(cx);
            right = c.box;
            double cy = c.y;
            for (int k = 0; k < n_extra; ++k) {
                cx += (car_span + MIN_DIST_TO_OBST + g.between(dlo, dmax));
                cy += g.gauss(0, 0.05, -0.1, 0.1);
                double yaw = g.gauss(yaw_mean, PI / 36, yaw_lo, yaw_hi);
                Ring e = car_box(cx, cy, yaw);
                if (g.uniform() < P_KEEP_EXTRA) extra.push_back(e);
            }
        }
    }
    double gap_r = ring_gap(dest_box, right);
    if (gap_r + gap_l < min_space || gap_r + gap_l > max_space || gap_l < MIN_DIST_TO_OBST || gap_r < MIN_DIST_TO_OBST) ok = false;
    obstacles.push_back(back); obstacles.push_back(left); obstacles.push_back(right);
    for (auto &e : extra) obstacles.push_back(e);
    for (auto &o : obstacles) if (rings_touch(o, dest_box)) ok = false;
    // far side of the aisle
    double top = -INFINITY;
    for (auto &o : obstacles) for (int i = 0; i < o.n; ++i) top = std::max(top, o.v[i].y);
    top += MIN_DIST_TO_OBST;
    std::vector<Ring> far;
    if (g.uniform() < 0.2) {
        double y0 = wall_dist + top + MIN_DIST_TO_OBST;
        Ring w;
        w.v[0] = {-half, y0}; w.v[1] = {half, y0}; w.v[2] = {half, y0 + 0.1}; w.v[3] = {-half, y0 + 0.1};
        far.push_back(w);
    } else {
        Ring zone;
        zone.v[0] = {-half, wall_dist + top}; zone.v[1] = {half, wall_dist + top}; zone.v[2] = {half, wall_dist + top + 8}; zone.v[3] = {-half, wall_dist + top + 8};
        for (int k = 0; k < N_OBSTACLE[level]; ++k) {
            double ox = g.between(-half + 2, half - 2), oy = g.between(wall_dist + top + 2, wall_dist + top + 6);
            Ring o = car_box(ox, oy, g.uniform() * PI * 2);
            for (int i = 0; i < 4; ++i) { o.v[i].x += 0.5 * g.uniform(); o.v[i].y += 0.5 * g.uniform(); }
            if (rings_touch(o, zone)) continue;
            bool clash = false;
            for (auto &f : far) if (rings_touch(o, f)) { clash = true; break; }
            if (!clash) far.push_back(o);
        }
    }
    for (auto &f : far) obstacles.push_back(f);
    // start pose: anywhere in the aisle that touches nothing
    double sx, sy, syaw;
    for (int guard = 0;; ++guard) {
        sx = g.between(-half / 2, half / 2);
        sy = g.between(top + 1, wall_dist + top - 1);
        syaw = g.gauss(0, PI / 6, -PI / 2, PI / 2);
        if (g.uniform() < 0.5) syaw += PI;
        Ring sb = car_box(sx, sy, syaw);
        bool free_ = !rings_touch(dest_box, sb);
        for (auto &o : obstacles) if (rings_touch(o, sb)) free_ = false;
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
    sc.obs = obstacles;
    sc.case_id = parallel ? 1 : 0;
    return true;
}

}  // namespace

extern "C" int hope_generate_scenes(int n, int level, uint64_t seed, int nthreads, double *h_start, double *h_dest, double *h_bounds,
                                    double *h_obs_xy, int32_t *h_nverts, int32_t *h_case_id) {
    if (n <= 0 || level < 0 || level > 2 || !h_start || !h_dest || !h_bounds || !h_obs_xy || !h_nverts) return HOPE_ERR_INVALID;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min(nthreads, n);
    std::vector<int> rc(nthreads, HOPE_OK);
    auto work = [&](int tid) {
        for (int i = tid; i < n; i += nthreads) {
            Rng g(seed + (uint64_t)i * 0x9e3779b97f4a7c15ull + 1);
            // ParkingMapNormal.reset :475-480 with case_id None
            bool bay = (g.uniform() > 0.5) && level != 2;
            Scene sc;
            int tries = 0;
            while (!attempt(g, level, !bay, sc)) if (++tries > 100000) { rc[tid] = HOPE_ERR_INVALID; return; }
            if ((int)sc.obs.size() > HOPE_MAX_OBS) { rc[tid] = HOPE_ERR_CAPACITY; return; }
            memcpy(h_start + 3 * i, sc.start, 24); memcpy(h_dest + 3 * i, sc.dest, 24);
            double *b = h_bounds + 4 * i;  // :486-489
            b[0] = floor(std::min(sc.start[0], sc.dest[0]) - 10); b[1] = ceil(std::max(sc.start[0], sc.dest[0]) + 10);
            b[2] = floor(std::min(sc.start[1], sc.dest[1]) - 10); b[3] = ceil(std::max(sc.start[1], sc.dest[1]) + 10);
            double *o = h_obs_xy + (size_t)i * HOPE_MAX_OBS * HOPE_MAX_VERTS * 2;
            int32_t *nv = h_nverts + (size_t)i * HOPE_MAX_OBS;
            memset(o, 0, sizeof(double) * HOPE_MAX_OBS * HOPE_MAX_VERTS * 2);
            memset(nv, 0, sizeof(int32_t) * HOPE_MAX_OBS);
            for (size_t k = 0; k < sc.obs.size(); ++k) {
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
