// render.cuh — k_render: the image observation without ever materialising the reference's
// 500 x 500 screen, its rotated copy or the 256 x 256 crop in HBM.  Included by hope_kernels.cu (namespace hope).
//
// Reference path (src/env/car_parking_base.py:301-350, src/env/observation_processor.py:6-23, env_wrapper.py:52-55):
//   _render               paint obstacles, start outline, dest box, vehicle box, last <= 20 trajectory boxes on a
//                         500 x 500 surface (pygame.draw.polygon on integer vertices)
//   _get_img_observation  pygame.transform.rotate(screen, degrees(heading)), blit centred, shift by the rotated
//                         offset of the vehicle-box centroid, crop the central 256 x 256
//   process_img           white -> black, cv2.resize to 64 x 64 (INTER_LINEAR), / 255
// cv2's 4x INTER_LINEAR reads only source pixels 4i+1 and 4i+2 of every axis, so an output pixel is the rounded
// mean of 4 crop pixels, and each crop pixel is ONE screen pixel through two integer shifts and the 16.16
// fixed-point rotation.
//   k_render_camera (thread / env): the composed integer map crop pixel -> screen pixel (Camera, 96 B per env).
//   k_render (CTA / env, 256 threads, ~44 KB of shared memory, 5 CTAs per SM; the four quadrants of the image go
//   through the same window one after the other, set-up and span table are built once):
//   1. set-up: the screen window the quadrant's 64 x 64 sample lattice can touch (<= 180 x 180 pixels) is kept as
//      one byte per pixel in shared memory; one thread per shape turns its ring into integer screen vertices and a
//      scan-conversion record (edges in draw_fillpoly's visiting order); the start-box outline (draw_line, Bresenham)
//      becomes 5 segment records whose pixel run on any row has a closed form.  Every KIND of shape is prepared by a warp of
//      its own (the branches run side by side), through ONE copy of the float64 set-up code;
//   2. paint, order-free: _render's painter's order is the order of increasing colour index, so "the later shape wins" is a
//      per-pixel MAXIMUM.  A window byte holds a thermometer code of the colour index (colour_code / traj_code below) and
//      spans are OR-ed in with shared-memory atomics: OR of thermometer codes = maximum, whichever thread paints whichever
//      (shape, row) pair first.  Every such pair is one work item — pygame's scan conversion restated literally per row —
//      and the items are dealt round all 256 threads (static shapes: by a running row offset; vehicle + trajectory boxes:
//      from a span table built once per image, 2a).  A byte has room for 8 thermometer levels and there are 25 colours: the
//      20 trajectory colours share one level plus a 3-bit thermometer of their GROUP of 5; the gather finds the newest box
//      of the newest group that covers the sample in the span table (at most 4 look-ups).
//      (Round 1 let one thread OWN one window row and replay the painter's order on it: the rows through the image centre
//      carry ~22 boxes, the others none, and half of all warp samples sat at the paint -> gather barrier; 6.5 ms per 65 536
//      images against 4.9 ms now.  The first order-free build was SLOWER (11 ms): 16.7 k SASS instructions, a third of
//      the stall samples "no instruction" — the L1.5 instruction cache holds 32 KB.  Out-of-line cold paths and the shared
//      set-up code brought it to 4.0 k instructions.)
//   3. gather: each thread resolves output pixels = 4 window bytes -> colour index -> palette -> (sum + 2) >> 2 and stores them
//      as uint8 [3][64][64], a warp writing one 32-byte sector per channel (the reference's float64 image is this / 255).
// HBM traffic per env-step: 12 288 B written + ~1.3 KB read (scene ring vertices, trajectory ring buffer; the other
// three quadrants hit L2).
#pragma once

namespace render {

constexpr int WIN = 500;          // configs.py:93-94 WIN_W, WIN_H
constexpr int OBS = 256;          // configs.py:89-90 OBS_W, OBS_H
constexpr int IMG = 64;           // OBS / downsample_rate (observation_processor.py:8)
constexpr int KSCALE = 12;        // configs.py:103 K
constexpr int TRAJ = 20;          // configs.py:86 TRAJ_RENDER_LEN
constexpr int QUAD = 32;          // output pixels per quadrant side (4 quadrants of the 64 x 64 image)
constexpr int PITCHW = 47;        // window words per row (188 pixels >= 125 sqrt(2) + 2 + alignment); odd, so the rows
                                  // owned by adjacent lanes start in different banks
constexpr int PITCH = PITCHW * 4;
constexpr int ROWS = 180;
constexpr int THREADS = 256;
constexpr int MAXSHAPES = MAXO + 3 + TRAJ;
constexpr int NDYN = 1 + TRAJ;    // vehicle box + trajectory boxes
constexpr int DROWS = 64;         // screen rows a vehicle-sized box can span: its diagonal is 5.07 m * K = 60.9 pixels
constexpr int NCOLOR = 5 + TRAJ;  // 0 background, 1 obstacle, 2 start outline, 3 dest, 4 vehicle, 5.. trajectory old -> new
// What a window byte holds: a thermometer code of the colour index, so that OR = "the later painted shape wins".  Indices 1..4:
// the low (index) bits set.  Trajectory boxes (20 colours, old -> new) set bits 0..4 plus a 3-bit thermometer of their GROUP
// (TGROUP consecutive boxes): the OR keeps the newest group exactly, and the gather finds the newest box of that group
// that covers the sample in the span table (<= TGROUP look-ups).
constexpr int TGROUP = 5;
static_assert((TRAJ + TGROUP - 1) / TGROUP - 1 <= 3, "group thermometer: 3 bits");
constexpr unsigned CODE_DYN = 0x10u;    // bit 4: a trajectory box covers the pixel
__host__ __device__ constexpr unsigned colour_code(int index) { return (1u << index) - 1u; }                                      // index 1..4
__host__ __device__ constexpr unsigned traj_code(int i) { return 0x1fu | (((1u << (i / TGROUP)) - 1u) << 5); }                    // i = 0 oldest
static_assert(TRAJ <= 32 && THREADS >= ROWS + 1 && MAXSHAPES <= THREADS - 2 - NCOLOR, "row owners + the probe thread; shape threads, palette threads and the camera thread are disjoint");

struct Palette { uint32_t rg[NCOLOR], b[NCOLOR]; };  // R | G << 16 and B: 16-bit lanes so four samples add without carry

struct Camera {
    // screen pixel of crop pixel (u, v): capture coordinates xc = u + cx0, yc = v + cy0, then
    // fx = a0 + a1 * xc + a2 * yc, fy = b0 + b1 * xc + b2 * yc in 16.16 fixed point (transform.c rotate())
    int cx0, cy0;          // crop -> capture offset
    int rx0, ry0;          // crop -> "rotate" surface offset (the 500 x 500 blit target)
    int nx, ny;            // capture size
    int a0, a1, a2, b0, b1, b2;
    int wx0, wy0, wx1, wy1;  // screen window of one quadrant (inclusive), wx0 aligned to 4; filled in by k_render
    double kbx, kby;       // coord_transform_matrix offsets
    int ulo, uhi, vlo, vhi;  // crop pixels that land on the rotated screen copy at all (the rest reads as background)
};
// what k_render derives per image quadrant: the screen window (inclusive, wx0 aligned to 4) its sample lattice can
// touch, and fast = 1 when every sample of the quadrant maps inside the screen (no per-sample checks)
struct QuadWindow { int wx0, wy0, wx1, wy1, fast; };
static_assert(sizeof(Camera) == 96, "one Camera per env in HBM");

struct Edge { short ylo, yhi, xlo, dx; int dy; float rdy; };  // non-horizontal edge, lower end first: x(y) = xlo + (y - ylo) dx / dy

struct Shape {   // one ring prepared for draw_fillpoly
    short miny, maxy, minx, maxx;
    short color, ne, nh, outline;      // color: the byte code painted (colour_code / traj_code); outline = 1: the width-1 start box, 0: filled
    Edge e[4];                         // in the order draw_fillpoly visits them (decides floor / ceil)
    short hy[4], hxa[4], hxb[4];       // horizontal edges strictly between miny and maxy (incl. the closing zero-length edge)
};

// One segment of the width-1 start box (draw.c draw_line).  kind 0: horizontal run or single point, 1: vertical,
// 2: x-major Bresenham, 3: y-major Bresenham (dx <= dy)
struct Seg { short ylo, yhi, x1, y1, xa, xb, dx, dy, sx, sy, err0, kind; float rdy; };

struct Smem {
    Camera cam;
    QuadWindow quad[4];
    Shape shapes[MAXSHAPES];
    int nshapes;
    uint32_t probe;                     // screen pixel (0, 0) in byte 0: rotate()'s background colour index
    int nstatic;                        // shapes [0, nstatic) are painted, [nstatic, nshapes) are the dynamic boxes, old -> new
    Seg seg[5];
    uint2 pal[NCOLOR];                  // (R | G << 16, B)
    short2 drange[NDYN];                // (miny, maxy) of dynamic box d: what the gather needs of its Shape
    short2 dyn[NDYN][DROWS];            // span of dynamic box d on screen row miny_d + r; x > y: nothing; x == DYN_DIRECT: evaluate
    alignas(16) uint32_t win[ROWS * PITCHW + 5];   // + gather slack, rounded to whole uint4 for the clear
};

// _coord_transform + pygame's (int) conversion of one world point
__device__ __forceinline__ void to_screen(const Camera &c, double x, double y, int &ix, int &iy) {
    ix = (int)((double)KSCALE * x + c.kbx);
    iy = (int)((double)KSCALE * y + c.kby);
}

// ring (nv world vertices) -> Shape.  pygame receives shapely's closed coordinate list: nv + 1 points, last = first.
__device__ void ring_shape(Shape &S, const Camera &c, const double *bx, const double *by, int nv, int color, int outline) {
    int px[5], py[5];
    int mny = 0x7fff, mxy = -0x8000, mnx = 0x7fff, mxx = -0x8000;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nv) {
            to_screen(c, bx[k], by[k], px[k], py[k]);
            mny = min(mny, py[k]); mxy = max(mxy, py[k]); mnx = min(mnx, px[k]); mxx = max(mxx, px[k]);
        }
    }
#pragma unroll
    for (int k = 1; k < 5; ++k) if (k == nv) { px[k] = px[0]; py[k] = py[0]; }
    const int n = nv + 1;
    S.miny = (short)mny; S.maxy = (short)mxy; S.minx = (short)mnx; S.maxx = (short)mxx;
    S.color = (short)color; S.outline = (short)outline;
    int ne = 0, nh = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {  // draw_fillpoly's edge loop: (p[i-1], p[i]) with p[-1] = p[n-1]
        if (i < n) {
            const int ip = i ? i - 1 : n - 1;
            const int y1 = py[ip], y2 = py[i];
            if (y1 != y2) {
                const bool up = y1 < y2;
                Edge E;
                E.ylo = (short)(up ? y1 : y2); E.yhi = (short)(up ? y2 : y1);
                E.xlo = (short)(up ? px[ip] : px[i]);
                E.dx = (short)(up ? px[i] - px[ip] : px[ip] - px[i]);
                E.dy = (int)E.yhi - (int)E.ylo; E.rdy = __frcp_rn((float)E.dy);
                if (ne < 4) S.e[ne] = E;
                ++ne;
            } else if (mny < y2 && y2 < mxy) {
                if (nh < 4) { S.hy[nh] = (short)y2; S.hxa[nh] = (short)px[i]; S.hxb[nh] = (short)px[ip]; }
                ++nh;
            }
        }
    }
    S.ne = (short)min(ne, 4); S.nh = (short)min(nh, 4);
}

// draw.c draw_line, one segment (x1, y1) -> (x2, y2).  The walk
//     err = (dx > dy ? dx : -dy) / 2;  loop: plot; e2 = err; if (e2 > -dx) { err -= dy; x += sx; } if (e2 < dy) { err += dx; y += sy; }
// advances x every iteration when dx > dy, with the i-th pixel on row offset ceil((i dy - err0) / dx); otherwise it
// advances y every iteration, with row offset j at column offset ceil((j dx + err0) / dy).  So the pixels of row offset k
// are i in (((k-1) dx + err0) / dy, (k dx + err0) / dy] (x-major) or the single column above (y-major) — checked
// exhaustively against the literal walk for |dx|, |dy| <= 40 (oracle/softraster.py line_pixels).
__device__ void make_seg(Seg &g, int x1, int y1, int x2, int y2) {
    g.ylo = (short)min(y1, y2); g.yhi = (short)max(y1, y2);
    g.x1 = (short)x1; g.y1 = (short)y1;
    g.xa = (short)min(x1, x2); g.xb = (short)max(x1, x2);
    const int dx = abs(x2 - x1), dy = abs(y2 - y1);
    g.dx = (short)dx; g.dy = (short)dy;
    g.sx = (short)(x1 < x2 ? 1 : -1); g.sy = (short)(y1 < y2 ? 1 : -1);
    g.kind = (short)(y1 == y2 ? 0 : (x1 == x2 ? 1 : (dx > dy ? 2 : 3)));
    g.err0 = (short)(dx > dy ? dx / 2 : -(dy / 2));
    g.rdy = dy ? __frcp_rn((float)dy) : 0.0f;
}

// inclusive run [x1, x2] of screen row `row` (byte offset of screen x is x - base), clipped to [cx0, cx1].  Pixels hold
// THERMOMETER codes (colour_code below) and are OR-ed in: a bytewise OR of thermometer codes is the bytewise maximum, and the
// painter's order of _render is the order of increasing colour index, so the result does not depend on which thread paints
// which (shape, row) pair first — any thread may paint any row (shared-memory atomics, no ownership, no ordering).
__device__ __forceinline__ void span(uint32_t *row, int base, int cx0, int cx1, int x1, int x2, uint32_t fill) {
    int a = max(min(x1, x2), cx0) - base, b = min(max(x1, x2), cx1) - base;
    if (b < a) return;
    const int wa = a >> 2, wb = b >> 2;
    const uint32_t ma = 0xffffffffu << (8 * (a & 3)), mb = 0xffffffffu >> (8 * (3 - (b & 3)));
    if (wa == wb) {
        atomicOr(&row[wa], fill & ma & mb);
        return;
    }
    atomicOr(&row[wa], fill & ma);
    for (int w = wa + 1; w < wb; ++w) atomicOr(&row[w], fill);
    atomicOr(&row[wb], fill & mb);
}

// exact floor / ceil of num / dy (dy > 0, |num| < 2^22): the reference's float division followed by floor / ceil
// is exact on these magnitudes, so integer arithmetic reproduces it
__device__ __forceinline__ int div_round(int num, int dy, float rdy, bool up) {
    int t = __float2int_rd(__int2float_rn(num) * rdy);
    int r = num - t * dy;
    if (r < 0) { --t; r += dy; } else if (r >= dy) { ++t; r -= dy; }
    return t + ((up && r != 0) ? 1 : 0);
}

constexpr short DYN_DIRECT = 0x7fff;


// The crossing list draw_fillpoly builds for row y of shape S (miny < maxy): values in visiting order, count returned.
__device__ __forceinline__ int row_crossings(const Shape &S, int y, int &x0, int &x1, int &x2, int &x3) {
    int cnt = 0;
    x0 = x1 = x2 = x3 = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (e < S.ne) {
            const Edge E = S.e[e];
            if ((y >= E.ylo && y < E.yhi) || (y == S.maxy && E.yhi == S.maxy)) {
                const int v = E.xlo + div_round((y - E.ylo) * E.dx, E.dy, E.rdy, cnt & 1);  // floor, ceil, floor, ceil
                if (cnt == 0) x0 = v; else if (cnt == 1) x1 = v; else if (cnt == 2) x2 = v; else x3 = v;
                ++cnt;
            }
        }
    }
    return cnt;
}

// Does the filled shape S paint screen pixel (sx, sy)?  Exact, straight from the scan conversion (used for the rare
// rows the span table cannot express, and for the background probe).
__device__ __noinline__ bool shape_covers(const Shape &S, int sx, int sy) {
    if (sy < S.miny || sy > S.maxy) return false;
    if (S.miny == S.maxy) return sx >= S.minx && sx <= S.maxx;
    int x0, x1, x2, x3;
    const int cnt = row_crossings(S, sy, x0, x1, x2, x3);
    bool hit = false;
    if (cnt == 2) hit = sx >= min(x0, x1) && sx <= max(x0, x1);
    else if (cnt == 3) {
        const int lo = min(x0, min(x1, x2)), hi = max(x0, max(x1, x2)), mid = x0 + x1 + x2 - lo - hi;
        hit = sx >= lo && sx <= mid;
    } else if (cnt == 4) {
        const int a = min(x0, x1), b = max(x0, x1), c = min(x2, x3), d = max(x2, x3);
        const int s0 = min(a, c), s3 = max(b, d), m1 = max(a, c), m2 = min(b, d);
        hit = (sx >= s0 && sx <= min(m1, m2)) || (sx >= max(m1, m2) && sx <= s3);
    }
    for (int k = 0; k < S.nh; ++k)
        if (S.hy[k] == sy && sx >= min(S.hxa[k], S.hxb[k]) && sx <= max(S.hxa[k], S.hxb[k])) hit = true;
    return hit;
}

// What one shape paints on screen row y (owned by the calling thread).
__device__ __noinline__ void paint_shape_row(const Smem &sm, const Shape &S, int y, uint32_t *row, int base, int cx0, int cx1) {
    {
        if (y < S.miny || y > S.maxy) return;
        const uint32_t fill = (uint32_t)S.color * 0x01010101u;
        if (S.outline) {  // lines(closed=True): 5 segments of the closed coordinate list
            for (int k = 0; k < 5; ++k) {
                const Seg g = sm.seg[k];
                if (y < g.ylo || y > g.yhi) continue;
                int xa = g.xa, xb = g.xb;  // kind 0
                if (g.kind == 1) { xa = xb = g.x1; }
                else if (g.kind >= 2) {
                    const int k_ = (y - g.y1) * g.sy;  // row offset along the walk, 0 .. dy
                    if (g.kind == 2) {
                        const int ilo = max(div_round((k_ - 1) * g.dx + g.err0, g.dy, g.rdy, false) + 1, 0);
                        const int ihi = min(div_round(k_ * g.dx + g.err0, g.dy, g.rdy, false), (int)g.dx);
                        xa = g.x1 + g.sx * ilo; xb = g.x1 + g.sx * ihi;
                    } else {
                        xa = xb = g.x1 + g.sx * div_round(k_ * g.dx + g.err0, g.dy, g.rdy, true);
                    }
                }
                span(row, base, cx0, cx1, xa, xb, fill);
            }
            return;
        }
        if (S.miny == S.maxy) { span(row, base, cx0, cx1, S.minx, S.maxx, fill); return; }  // one pixel high
        int x0, x1, x2, x3;
        const int cnt = row_crossings(S, y, x0, x1, x2, x3);
        if (cnt == 2 || cnt == 3) {
            if (cnt == 3) {  // sorted, the first two form the pair (the third has no partner)
                const int lo = min(x0, min(x1, x2)), hi = max(x0, max(x1, x2));
                const int mid = x0 + x1 + x2 - lo - hi;
                x0 = lo; x1 = mid;
            }
            span(row, base, cx0, cx1, x0, x1, fill);
        } else if (cnt == 4) {
            int a = min(x0, x1), b = max(x0, x1), c = min(x2, x3), d = max(x2, x3);
            const int s0 = min(a, c), s3 = max(b, d), m1 = max(a, c), m2 = min(b, d);
            span(row, base, cx0, cx1, s0, min(m1, m2), fill);
            span(row, base, cx0, cx1, max(m1, m2), s3, fill);
        }
        for (int k = 0; k < S.nh; ++k)
            if (S.hy[k] == y) span(row, base, cx0, cx1, S.hxa[k], S.hxb[k], fill);
    }
}

static_assert(MAXO != 16 || sizeof(Smem) <= 44400, "5 CTAs per SM (227 KB, 1 KB reserved per CTA)");

}  // namespace render

#ifndef HOPE_RENDER_HOST_TEST  // tests/render_host_harness.cpp compiles the scan-conversion helpers above with g++

// One thread per env: the camera of this step (see Camera).
__global__ void __launch_bounds__(128) k_render_camera(int n, Pool pool, EnvState st, hope_params par, render::Camera *__restrict__ cams) {
    using namespace render;
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const double *meta = pool.meta + (size_t)st.scene[env] * META;
    const double x = st.pose[3 * env], y = st.pose[3 * env + 1], h = st.pose[3 * env + 2];
    const double ch = st.cs[2 * env], sh = st.cs[2 * env + 1];
    Camera c;
    c.kbx = 0.5 * (WIN - KSCALE * (meta[M_BOUNDS + 1] + meta[M_BOUNDS]));  // car_parking_base.py:143-144
    c.kby = 0.5 * (WIN - KSCALE * (meta[M_BOUNDS + 3] + meta[M_BOUNDS + 2]));
    // centroid of the vehicle ring (GEOS lineal centroid, see oracle/geom.py), then _coord_transform (:328)
    double qx[5], qy[5];
    vehicle_box(x, y, ch, sh, par.box_x, par.box_y, qx, qy);
    qx[4] = qx[0]; qy[4] = qy[0];
    double tot = 0.0, sx = 0.0, sy = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double ddx = qx[k] - qx[k + 1], ddy = qy[k] - qy[k + 1];
        const double seg = sqrt(ddx * ddx + ddy * ddy);
        if (seg == 0.0) continue;
        tot += seg;
        sx += seg * ((qx[k] + qx[k + 1]) / 2);
        sy += seg * ((qy[k] + qy[k + 1]) / 2);
    }
    const double vcx = KSCALE * (sx / tot) + c.kbx, vcy = KSCALE * (sy / tot) + c.kby;
    const double ddx = (vcx - WIN / 2) * ch + (vcy - WIN / 2) * sh;   // :329-332
    const double ddy = -(vcx - WIN / 2) * sh + (vcy - WIN / 2) * ch;
    const int ox = (int)(-ddx), oy = (int)(-ddy);
    // pygame.transform.rotate: the angle is a C float of the degrees
    const float angle = (float)(h * (180.0 / HOPE_PI));
    if (fmod((double)angle, 90.0) == 0.0) {  // rotate90 path: exact quarter turns of the 500 x 500 screen
        int q = ((int)angle / 90) % 4;
        if (q < 0) q += 4;
        c.nx = WIN; c.ny = WIN;
        const int one = 1 << 16, last = (WIN - 1) << 16;
        if (q == 0) { c.a0 = 0; c.a1 = one; c.a2 = 0; c.b0 = 0; c.b1 = 0; c.b2 = one; }
        else if (q == 1) { c.a0 = last; c.a1 = 0; c.a2 = -one; c.b0 = 0; c.b1 = one; c.b2 = 0; }
        else if (q == 2) { c.a0 = last; c.a1 = -one; c.a2 = 0; c.b0 = last; c.b1 = 0; c.b2 = -one; }
        else { c.a0 = 0; c.a1 = 0; c.a2 = one; c.b0 = last; c.b1 = -one; c.b2 = 0; }
    } else {
        const double rad = (double)angle * .01745329251994329;
        double sa, ca;
        sincos(rad, &sa, &ca);
        const double cx = ca * WIN, cy = ca * WIN, sxx = sa * WIN, syy = sa * WIN;
        c.nx = (int)fmax(fmax(fmax(fabs(cx + syy), fabs(cx - syy)), fabs(-cx + syy)), fabs(-cx - syy));
        c.ny = (int)fmax(fmax(fmax(fabs(sxx + cy), fabs(sxx - cy)), fabs(-sxx + cy)), fabs(-sxx - cy));
        const int cyc = c.ny / 2, xd = (WIN - c.nx) << 15, yd = (WIN - c.ny) << 15;
        const int isin = (int)(sa * 65536), icos = (int)(ca * 65536);
        const int ax = (c.nx << 15) - (int)(ca * ((c.nx - 1) << 15));
        const int ay = (c.ny << 15) - (int)(sa * ((c.nx - 1) << 15));
        c.a0 = ax + isin * cyc + xd; c.a1 = icos; c.a2 = -isin;
        c.b0 = ay - icos * cyc + yd; c.b1 = isin; c.b2 = icos;
    }
    // crop pixel (u, v) = observation pixel (122 + u, 122 + v) = rotate-surface pixel shifted by the blit
    // offset = capture pixel shifted by the centred blit (Rect.center setter: x = cx - w / 2)
    const int crop0 = (WIN - OBS) / 2;
    c.rx0 = crop0 - ox; c.ry0 = crop0 - oy;
    c.cx0 = c.rx0 - (WIN / 2 - (c.nx >> 1)); c.cy0 = c.ry0 - (WIN / 2 - (c.ny >> 1));
    c.wx0 = c.wy0 = c.wx1 = c.wy1 = 0; c.ulo = c.uhi = c.vlo = c.vhi = 0;
    cams[env] = c;
}

// One CTA per env.  Set-up and the span table of the dynamic boxes are built once, then the four image quadrants are
// painted and gathered one after the other through the same shared-memory window.
// traj: [N][20][4] ring buffer (x, y, cos h, sin h) of Vehicle.trajectory's tail, traj_n: [N] its length (see k_advance)
__global__ void __launch_bounds__(render::THREADS, 5)
k_render(int n, Pool pool, EnvState st, const render::Camera *__restrict__ cams, hope_params par, render::Palette pal,
         uint8_t *__restrict__ img) {
    using namespace render;
    extern __shared__ __align__(16) unsigned char render_smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(render_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int env = blockIdx.x;
    const int sid = st.scene[env];
    const double *meta = pool.meta + (size_t)sid * META;
    // ---------------------------------------------------------------- 1. set-up, one phase: camera + the four windows
    // (one otherwise idle thread), shapes (one thread each; they only need the two screen offsets), palette
    if (tid == THREADS - 2) {
        Camera c = cams[env];
        // crop pixels inside the 500 x 500 blit target AND inside the rotated copy (both axis-aligned in u, v)
        c.ulo = max(-c.rx0, -c.cx0); c.uhi = min(WIN - 1 - c.rx0, c.nx - 1 - c.cx0);
        c.vlo = max(-c.ry0, -c.cy0); c.vhi = min(WIN - 1 - c.ry0, c.ny - 1 - c.cy0);
        sm.cam = c;
        const int fmax_ = (WIN << 16) - 1;
        for (int quad = 0; quad < 4; ++quad) {
            const int u0 = (quad & 1) * (OBS / 2), v0 = (quad >> 1) * (OBS / 2);
            // screen window touched by the sample lattice u, v in {4i+1, 4i+2} of this quadrant: the map is affine, so
            // the corners bound it
            int fx0 = 0x7fffffff, fx1 = -0x7fffffff - 1, fy0 = 0x7fffffff, fy1 = -0x7fffffff - 1;
            for (int k = 0; k < 4; ++k) {
                const int xc = u0 + ((k & 1) ? OBS / 2 - 2 : 1) + c.cx0, yc = v0 + ((k & 2) ? OBS / 2 - 2 : 1) + c.cy0;
                const int fx = c.a0 + c.a1 * xc + c.a2 * yc, fy = c.b0 + c.b1 * xc + c.b2 * yc;
                fx0 = min(fx0, fx); fx1 = max(fx1, fx); fy0 = min(fy0, fy); fy1 = max(fy1, fy);
            }
            QuadWindow w;
            w.wx0 = min(max(fx0 >> 16, 0), WIN - 1) & ~3; w.wx1 = min(min(max(fx1 >> 16, 0), WIN - 1), w.wx0 + PITCH - 1);
            w.wy0 = min(max(fy0 >> 16, 0), WIN - 1); w.wy1 = min(min(max(fy1 >> 16, 0), WIN - 1), w.wy0 + ROWS - 1);
            w.fast = (u0 + 1 >= c.ulo && u0 + OBS / 2 - 2 <= c.uhi && v0 + 1 >= c.vlo && v0 + OBS / 2 - 2 <= c.vhi &&
                      fx0 >= 0 && fy0 >= 0 && fx1 <= fmax_ && fy1 <= fmax_) ? 1 : 0;
            sm.quad[quad] = w;
        }
        sm.probe = 0u;
    }
    if (tid >= THREADS - 2 - NCOLOR && tid < THREADS - 2) sm.pal[tid - (THREADS - 2 - NCOLOR)] = make_uint2(pal.rg[tid - (THREADS - 2 - NCOLOR)], pal.b[tid - (THREADS - 2 - NCOLOR)]);
    {
        const int nobs = pool.nobs[sid];
        const int tn = st.traj_n[env];
        const int ntraj = tn > 1 ? min(tn, TRAJ) : 0;
        const int total = nobs + 3 + ntraj;
        if (tid == 0) { sm.nshapes = total; sm.nstatic = nobs + 2; }
        // which shape this thread prepares.  With <= 32 obstacles every KIND of shape gets a warp of its own (obstacles, trajectory
        // boxes, start, dest, vehicle): the five branches below then run side by side on the SM's schedulers instead of one after
        // the other inside warp 0, and this phase, which the other warps sit out at the barrier, is as long as its longest branch
        int s = tid < total ? tid : -1;
        if constexpr (MAXO <= 32) {
            const int w = tid >> 5;
            s = -1;
            if (w == 0) { if (lane < nobs) s = lane; }
            else if (w == 1) { if (lane < ntraj) s = nobs + 3 + lane; }
            else if (w <= 4 && lane == 0) s = nobs + (w - 2);
        }
        if (s >= 0) {
            Shape &S = sm.shapes[s];
            double bx[4], by[4];
            Camera cam;  // only the screen offsets are read here
            cam.kbx = cams[env].kbx; cam.kby = cams[env].kby;
            // the branches only fetch what describes the shape; the box corners, the scan-conversion record and the outline
            // segments are built once below (one copy of that code in the instruction cache, no divergence inside it)
            int nv = 4, code, outline = 0;
            bool is_box = true, skip = false;
            double x = 0.0, y = 0.0, c = 1.0, sn = 0.0;
            if (s < nobs) {  // :303-305 obstacles
                nv = pool.nv[(size_t)sid * MAXO + s];
                const double2 *v = reinterpret_cast<const double2 *>(pool.obs) + ((size_t)sid * MAXO + s) * MAXV;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < nv) { const double2 p = __ldg(v + k); bx[k] = p.x; by[k] = p.y; }
                code = (int)colour_code(1); is_box = false;
            } else if (s == nobs) {  // :307-308 start box, width = 1: lines(closed=True) over the 5 coordinates
                sincos(meta[M_START + 2], &sn, &c);
                x = meta[M_START]; y = meta[M_START + 1];
                code = (int)colour_code(2); outline = 1;
            } else if (s == nobs + 1) {  // :309-310 dest box
#pragma unroll
                for (int k = 0; k < 4; ++k) { bx[k] = meta[M_DBX + k]; by[k] = meta[M_DBY + k]; }
                code = (int)colour_code(3); is_box = false;
            } else if (s == nobs + 2) {  // :312-313 vehicle
                x = st.pose[3 * env]; y = st.pose[3 * env + 1]; c = st.cs[2 * env]; sn = st.cs[2 * env + 1];
                code = (int)colour_code(4);
                if (ntraj > 0) {  // the newest trajectory box is painted later over the very same pixels: skip this one
                    const double2 *p = reinterpret_cast<const double2 *>(st.traj) + ((size_t)env * TRAJ + (tn - 1) % TRAJ) * 2;
                    const double2 xy = p[0], cs = p[1];
                    skip = xy.x == x && xy.y == y && cs.x == c && cs.y == sn;
                }
            } else {  // :315-319 trajectory[-(ntraj - i)], colour TRAJ_COLORS[-(ntraj - i)]
                const int i = s - (nobs + 3), back = ntraj - i;  // back = 1: newest
                const double2 *p = reinterpret_cast<const double2 *>(st.traj) + ((size_t)env * TRAJ + (tn - back) % TRAJ) * 2;
                const double2 xy = p[0], cs = p[1];
                x = xy.x; y = xy.y; c = cs.x; sn = cs.y;
                code = (int)traj_code(i);  // palette index 5 + TRAJ - back
            }
            if (is_box) vehicle_box(x, y, c, sn, par.box_x, par.box_y, bx, by);
            ring_shape(S, cam, bx, by, nv, code, outline);
            if (skip) { S.miny = 1; S.maxy = 0; }
            if (outline) {
                int px[4], py[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) to_screen(cam, bx[q], by[q], px[q], py[q]);
#pragma unroll 1
                for (int k = 0; k < 5; ++k) {  // (p4 = p0) -> p0: a single pixel
                    const int k0 = k & 3, k1 = k == 4 ? 0 : (k + 1) & 3;
                    make_seg(sm.seg[k], px[k0], py[k0], px[k1], py[k1]);
                }
            }
            if (s >= nobs + 2) sm.drange[s - (nobs + 2)] = make_short2(S.miny, S.maxy);
        }
    }
    __syncthreads();
    const Camera &cam = sm.cam;
    const int nstatic = sm.nstatic, ndyn = sm.nshapes - nstatic;
    // ---------------------------------------------------------------- 2a. span table of the dynamic boxes: one
    // (box, row) pair per thread and pass, all threads, once per env; the window is cleared for the first quadrant
    for (int idx = tid; idx < ndyn * DROWS; idx += THREADS) {
        const int d = idx / DROWS, r = idx - d * DROWS;
        const Shape &S = sm.shapes[nstatic + d];
        const int y = S.miny + r;
        short2 e = make_short2(1, 0);  // nothing on this row
        if (y <= S.maxy && y >= 0 && y < WIN) {
            if (S.miny == S.maxy) e = make_short2(S.minx, S.maxx);
            else {
                int x0, x1, x2, x3;
                const int cnt = row_crossings(S, y, x0, x1, x2, x3);
                if (cnt == 2) e = make_short2((short)min(x0, x1), (short)max(x0, x1));
                else if (cnt == 3) {
                    const int lo = min(x0, min(x1, x2)), hi = max(x0, max(x1, x2));
                    e = make_short2((short)lo, (short)(x0 + x1 + x2 - lo - hi));
                } else if (cnt == 4) e = make_short2(DYN_DIRECT, 0);  // two runs: left to paint_shape_row / shape_covers
                for (int k = 0; k < S.nh; ++k)  // a horizontal edge on this row (the closing zero-length edge, usually): one more run
                    if (S.hy[k] == y && e.x != DYN_DIRECT) {
                        const int c = min(S.hxa[k], S.hxb[k]), dd = max(S.hxa[k], S.hxb[k]);
                        if (e.x > e.y) e = make_short2((short)c, (short)dd);
                        else if (c <= e.y + 1 && dd >= e.x - 1) e = make_short2((short)min((int)e.x, c), (short)max((int)e.y, dd));  // touches the run: one run
                        else e = make_short2(DYN_DIRECT, 0);
                    }
            }
        }
        sm.dyn[d][r] = e;
    }
    constexpr int WINQ = (int)(sizeof(sm.win) / 16);
    for (int k = tid; k < WINQ; k += THREADS) reinterpret_cast<uint4 *>(sm.win)[k] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == THREADS - 1) {  // screen pixel (0, 0), rotate()'s background colour, as a 1-pixel "row"; kept as a palette index
        for (int s = 0; s < nstatic; ++s) {
            const Shape &S = sm.shapes[s];
            if (S.miny <= 0 && S.maxy >= 0 && S.minx <= 0 && S.maxx >= 0) paint_shape_row(sm, S, 0, &sm.probe, 0, 0, 0);
        }
        unsigned idx = __popc(sm.probe & 0xffu);
        for (int d = 0; d < ndyn; ++d)
            if (sm.shapes[nstatic + d].miny <= 0 && sm.shapes[nstatic + d].minx <= 0 && shape_covers(sm.shapes[nstatic + d], 0, 0)) idx = d == 0 ? 4u : (unsigned)(5 + TRAJ - ndyn + d);  // trajectory i = d - 1 of ndyn - 1
        sm.probe = idx;
    }
    __syncthreads();
    const unsigned bgidx = sm.probe & 0xffu;
    const int ntraj = ndyn - 1;
    // palette index of a window byte at screen pixel (sx, sy): thermometer codes 0..4 by bit count, trajectory bytes through resolve_traj
    // A trajectory byte names the newest GROUP of boxes that covers the pixel; the newest box of that group whose stored run
    // contains sx is the one on top (and if none of the newer ones does, it is the group's oldest: one of them set the bit).
    auto resolve = [&](unsigned code, int sx, int sy) -> unsigned {
        if (code < CODE_DYN) return (unsigned)__popc(code);
        const int g0 = __popc(code >> 5) * TGROUP;
        int i = min(g0 + TGROUP - 1, ntraj - 1);
#pragma unroll 1
        for (; i > g0; --i) {
            const short2 rg = sm.drange[1 + i];
            const int r = sy - rg.x;
            if ((unsigned)r >= (unsigned)DROWS || sy > rg.y) continue;
            const short2 e = sm.dyn[1 + i][r];
            if (sx >= e.x && sx <= e.y) break;
            if (e.x == DYN_DIRECT && shape_covers(sm.shapes[nstatic + 1 + i], sx, sy)) break;
        }
        return (unsigned)(5 + TRAJ - ntraj + i);
    };
#pragma unroll 1
    for (int quad = 0; quad < 4; ++quad) {
        const QuadWindow win = sm.quad[quad];
        const int u0 = (quad & 1) * (OBS / 2), v0 = (quad >> 1) * (OBS / 2);  // crop origin of this quadrant
        // ------------------------------------------------------------ 2b. paint: every (shape, window row) pair is one work
        // item, any thread paints any item (span() ORs thermometer codes).  Static shapes: their rows are dealt round the threads;
        // dynamic boxes: their stored runs.
        int first = 0;  // rows handed out so far (mod 256): shape after shape the items go round the threads, so every thread
                        // gets floor or ceil of (all rows of all shapes) / 256 of them
        for (int sb = 0; sb < nstatic; sb += 32) {  // each lane tests one shape against the window, the warp walks the hits
            bool touch = false;
            if (sb + lane < nstatic) {
                const Shape &S = sm.shapes[sb + lane];
                touch = S.maxy >= win.wy0 && S.miny <= win.wy1 && S.maxx >= win.wx0 && S.minx <= win.wx1;
            }
            unsigned m = __ballot_sync(HOPE_FULL_MASK, touch);
            while (m) {
                const int s = sb + __ffs(m) - 1;
                m &= m - 1;
                const Shape &S = sm.shapes[s];
                const int lo = max((int)S.miny, win.wy0), cnt = min((int)S.maxy, win.wy1) - lo + 1;
                const int t = (tid - first) & (THREADS - 1);
                first += cnt;
                if (t < cnt) paint_shape_row(sm, S, lo + t, sm.win + (lo + t - win.wy0) * PITCHW, win.wx0, win.wx0, win.wx1);
            }
        }
        for (int idx = tid; idx < ndyn * DROWS; idx += THREADS) {
            const int d = idx / DROWS, r = idx - d * DROWS;
            const Shape &S = sm.shapes[nstatic + d];
            const int y = S.miny + r;
            if (y > S.maxy || y < win.wy0 || y > win.wy1) continue;
            uint32_t *row = sm.win + (y - win.wy0) * PITCHW;
            const short2 e = sm.dyn[d][r];
            if (e.x == DYN_DIRECT) { paint_shape_row(sm, S, y, row, win.wx0, win.wx0, win.wx1); continue; }
            const uint32_t fill = (uint32_t)S.color * 0x01010101u;
            if (e.x <= e.y) span(row, win.wx0, win.wx0, win.wx1, e.x, e.y, fill);
        }
        __syncthreads();
        // ------------------------------------------------------------ 3. gather 32 x 32 x (2 x 2 samples)
        {
            const unsigned char *win8 = reinterpret_cast<const unsigned char *>(sm.win);
            const int a0 = cam.a0, a1 = cam.a1, a2 = cam.a2, b0 = cam.b0, b1 = cam.b1, b2 = cam.b2;
            const int woff = win.wy0 * PITCH + win.wx0;
            const int i = tid & (QUAD - 1);
            const int ub = u0 + 4 * i + 1, xc = ub + cam.cx0;
            uint8_t *out = img + (size_t)env * 3 * IMG * IMG + (quad >> 1) * QUAD * IMG + (quad & 1) * QUAD + i;
            if (win.fast) {
#pragma unroll 1
                for (int r = 0; r < QUAD * QUAD / THREADS; ++r) {
                    const int j = (tid >> 5) + r * (THREADS / QUAD);
                    const int yc = v0 + 4 * j + 1 + cam.cy0;
                    const int fxb = a0 + a1 * xc + a2 * yc, fyb = b0 + b1 * xc + b2 * yc;
                    uint32_t srg = 0u, sb = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int fx = fxb + ((k & 1) ? a1 : 0) + ((k >> 1) ? a2 : 0), fy = fyb + ((k & 1) ? b1 : 0) + ((k >> 1) ? b2 : 0);
                        const uint2 c = sm.pal[resolve(win8[(fy >> 16) * PITCH + (fx >> 16) - woff], fx >> 16, fy >> 16)];
                        srg += c.x; sb += c.y;
                    }
                    srg = ((srg + 0x00020002u) >> 2) & 0x00ff00ffu;  // (a + b + c + d + 2) >> 2 per 16-bit lane
                    sb = (sb + 2u) >> 2;
                    uint8_t *o = out + j * IMG;  // a warp writes 32 consecutive bytes per channel: one full sector each
                    o[0] = (uint8_t)(srg & 0xffu); o[IMG * IMG] = (uint8_t)(srg >> 16); o[2 * IMG * IMG] = (uint8_t)sb;
                }
            } else {
                const unsigned xmaxv = (WIN << 16) - 1;
                const int ulo = cam.ulo, uhi = cam.uhi, vlo = cam.vlo, vhi = cam.vhi;
                for (int r = 0; r < QUAD * QUAD / THREADS; ++r) {
                    const int j = (tid >> 5) + r * (THREADS / QUAD);
                    const int vb = v0 + 4 * j + 1, yc = vb + cam.cy0;
                    const int fxb = a0 + a1 * xc + a2 * yc, fyb = b0 + b1 * xc + b2 * yc;
                    uint32_t srg = 0u, sb = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int u = ub + (k & 1), v = vb + (k >> 1);
                        unsigned idx = 0u;  // background: white -> black, and the untouched (black) blit target
                        if (u >= ulo && u <= uhi && v >= vlo && v <= vhi) {
                            const int fx = fxb + ((k & 1) ? a1 : 0) + ((k >> 1) ? a2 : 0), fy = fyb + ((k & 1) ? b1 : 0) + ((k >> 1) ? b2 : 0);
                            if ((unsigned)fx > xmaxv || (unsigned)fy > xmaxv) idx = bgidx;  // negative wraps to a huge unsigned
                            else {
                                const unsigned off = (unsigned)((fy >> 16) * PITCH + (fx >> 16) - woff);
                                if (off < (unsigned)(ROWS * PITCH)) idx = resolve(win8[off], fx >> 16, fy >> 16);
                            }
                        }
                        const uint2 c = sm.pal[idx];
                        srg += c.x; sb += c.y;
                    }
                    srg = ((srg + 0x00020002u) >> 2) & 0x00ff00ffu;
                    sb = (sb + 2u) >> 2;
                    uint8_t *o = out + j * IMG;
                    o[0] = (uint8_t)(srg & 0xffu); o[IMG * IMG] = (uint8_t)(srg >> 16); o[2 * IMG * IMG] = (uint8_t)sb;
                }
            }
        }
        if (quad < 3) {  // the window is cleared and repainted for the next quadrant
            __syncthreads();
            for (int k = tid; k < WINQ; k += THREADS) reinterpret_cast<uint4 *>(sm.win)[k] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }
    }
}

#endif  // HOPE_RENDER_HOST_TEST
