// render.cuh — the image observation without ever materialising the reference's rotated screen copy or
// the 256 x 256 crop; of the 500 x 500 screen only the part that is constant during an episode is kept (2 bits per pixel).  Included by hope_kernels.cu (namespace hope).
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
//   k_render_static (CTA / env, returns at once unless the env's scene changed): obstacles, start outline and dest box do not
//   change during an episode, so they are painted ONCE per (env, scene) into HBM as 2 bits per pixel (colour index 0..3), 128
//   bytes per screen row, 64 000 bytes per env (4.2 GB for 65 536 envs).  Painting is order-free: _render's painter's order is
//   the order of increasing colour index, so "the later shape wins" is a per-pixel MAXIMUM; a tile byte holds a thermometer code
//   of the colour index and spans are OR-ed in with shared-memory atomics, every (shape, row) pair one work item dealt round
//   the 256 threads (pygame's scan conversion restated literally per row; the start-box outline = 5 Bresenham segments whose
//   pixel run on any row has a closed form).
//   k_render (CTA / env, 256 threads, 54 KB of shared memory, 4 CTAs per SM), every step:
//   1. set-up: the screen window the 128 x 128 sample lattice can touch (<= 360 x 360 pixels, rotated view); one thread per
//      dynamic box (vehicle + the last <= 20 trajectory boxes) prepares its scan-conversion record;
//   2. the window's rows of the cached static screen are fetched with cp.async.bulk (one copy per row, <= 96 B, completion
//      counted by one mbarrier) and land under steps 3 and 4.  They go through a 26.7 KB buffer one HALF of the image at a time
//      (the rows of output rows 32-63 are requested when the gather of rows 0-31 is done): the whole view would take 40 KB in the
//      worst rotation, and the difference is a fourth resident CTA per SM;
//   3. span table: the run every dynamic box paints on each of its <= 64 screen rows.  A trajectory box stays where it is for the
//      20 steps it is drawn (only the camera moves), so its runs are computed ONCE, the step it joins the trail, and kept in HBM
//      beside the pose they belong to (6.4 KB per env: one record per slot of the trajectory ring buffer).  All 20 records arrive
//      by one 5 KB bulk copy requested before anything else; a record is used iff its pose and screen offsets are the ones this
//      step would compute from, else the runs are computed (one (box, row) pair per thread and pass) and the record rewritten;
//   4. dynamic layer: windows of one NIBBLE per screen pixel over the boxes' bounding box (cut to the view): a thread owns a window
//      row and writes its boxes' runs on it old -> new, whole words between the ends.  The older and the newer half of the boxes
//      have a window each (2 x 12 KB; the gather asks the newer one first): the two halves are painted side by side by the two
//      half-warps, a row owner walks at most 11 boxes, and 11 codes fit the nibble.  A trail whose windows exceed their 12 KB (more
//      than ~ 140 x 140 pixels per half inside the view) is resolved per lattice sample instead: newest box first, every thread
//      tests ITS samples of the box's lattice rectangle against the span table;
//   5. gather: a warp resolves 8 x 4 output pixels per round; an output pixel = 4 samples; a sample = the dynamic layer's box if any,
//      else the 2-bit static pixel of the staged rows; palette sums in one packed word (10-bit lanes), rounded mean, uint8
//      [3][64][64] (the reference's float64 image is this / 255).
// HBM traffic per env-step: 12 288 B written + 20-41 KB of static rows + 6.4 KB of span records read + ~1.3 KB of pose /
// trajectory (measured: 48 KB read, 11.5 KB written per image).
// History (profiles/r02_ncu_k_render_history.txt): round 1 painted the whole window per step with one thread per window row
// (6.5 ms per 65 536 images); order-free painting of four quadrant windows 4.9 ms; this design 2.7 ms.
#pragma once

namespace render {

constexpr int WIN = 500;          // configs.py:93-94 WIN_W, WIN_H
constexpr int OBS = 256;          // configs.py:89-90 OBS_W, OBS_H
constexpr int IMG = 64;           // OBS / downsample_rate (observation_processor.py:8)
constexpr int KSCALE = 12;        // configs.py:103 K
constexpr int TRAJ = 20;          // configs.py:86 TRAJ_RENDER_LEN
constexpr int THREADS = 256;
constexpr int NDYN = 1 + TRAJ;    // vehicle box + trajectory boxes
constexpr int DROWS = 64;         // screen rows a vehicle-sized box can span: its diagonal is 5.07 m * K = 60.9 pixels
constexpr int NCOLOR = 5 + TRAJ;  // 0 background, 1 obstacle, 2 start outline, 3 dest, 4 vehicle, 5.. trajectory old -> new
// What a tile byte of k_render_static holds: a thermometer code of the colour index (1..3: the low `index` bits set), so that
// OR = "the later painted shape wins".
__host__ __device__ constexpr unsigned colour_code(int index) { return (1u << index) - 1u; }
static_assert(TRAJ <= 32 && MAXO + 2 <= THREADS && 33 <= THREADS - 2 - NCOLOR, "one thread per dynamic box; shape threads, palette threads and the camera thread are disjoint");

struct Palette { uint32_t rg[NCOLOR], b[NCOLOR]; };  // R | G << 16 and B (k_render repacks them into 10-bit lanes)

struct Camera {
    // screen pixel of crop pixel (u, v): capture coordinates xc = u + cx0, yc = v + cy0, then
    // fx = a0 + a1 * xc + a2 * yc, fy = b0 + b1 * xc + b2 * yc in 16.16 fixed point (transform.c rotate())
    int cx0, cy0;          // crop -> capture offset
    int rx0, ry0;          // crop -> "rotate" surface offset (the 500 x 500 blit target)
    int nx, ny;            // capture size
    int a0, a1, a2, b0, b1, b2;
    int wx0, wy0, wx1, wy1;  // on-screen part of the sample lattice's bounding box (inclusive); filled in by k_render
    double kbx, kby;       // coord_transform_matrix offsets
    int ulo, uhi, vlo, vhi;  // crop pixels that land on the rotated screen copy at all (the rest reads as background)
};
static_assert(sizeof(Camera) == 96, "one Camera per env in HBM");

struct Edge { short ylo, yhi, xlo, dx; int dy; float rdy; };  // non-horizontal edge, lower end first: x(y) = xlo + (y - ylo) dx / dy

struct Shape {   // one ring prepared for draw_fillpoly
    short miny, maxy, minx, maxx;
    short color, ne, nh, outline;      // color: the byte code painted (colour_code; unused for the dynamic boxes); outline = 1: the width-1 start box, 0: filled
    Edge e[4];                         // in the order draw_fillpoly visits them (decides floor / ceil)
    short hy[4], hxa[4], hxb[4];       // horizontal edges strictly between miny and maxy (incl. the closing zero-length edge)
};

// One segment of the width-1 start box (draw.c draw_line).  kind 0: horizontal run or single point, 1: vertical,
// 2: x-major Bresenham, 3: y-major Bresenham (dx <= dy)
struct Seg { short ylo, yhi, x1, y1, xa, xb, dx, dy, sx, sy, err0, kind; float rdy; };

// The static part of the screen (obstacles, start outline, dest box) does not change during an episode: k_render_static paints
// it ONCE per (env, scene) into HBM as 2 bits per pixel (the colour index 0..3), 128 bytes per screen row (500 pixels + padding, so
// that every row and every 16-byte chunk of it is aligned for bulk copies): 64 000 bytes per env.  k_render stages the rows its
// sample lattice can touch into shared memory (cp.async.bulk, one copy per row, one mbarrier) and reads the pixels from there.
constexpr int SCREEN_PITCH = 128;                  // bytes per screen row
constexpr int SCREEN_BYTES = WIN * SCREEN_PITCH;   // per env
constexpr int TROWS = 100;                         // screen rows per shared-memory tile of k_render_static
constexpr int TPITCHW = SCREEN_PITCH;              // words per tile row: one thermometer byte per pixel while painting, 512 pixels
constexpr int TILEW = TROWS * TPITCHW;             // words per tile
static_assert(WIN <= 4 * SCREEN_PITCH && SCREEN_PITCH % 16 == 0 && WIN % TROWS == 0, "padded rows, whole tiles");

struct Smem {                           // k_render_static
    Shape shapes[MAXO + 2];             // obstacles, start outline, dest box: the painter's order
    Seg seg[5];
    int nstatic;
    alignas(16) uint32_t tile[TILEW];
};

// k_render's sample lattice: cv2's 4x INTER_LINEAR reads crop pixels 4i+1, 4i+2 of either axis = 128 x 128 samples,
// lattice index a <-> crop pixel 2a + 1 - (a & 1)
constexpr int LAT = 2 * IMG;
// The static rows are staged per HALF of the image (output rows 0-31, then 32-63, through the same buffer): half the lattice touches
// at most 253 |sin| + 125 |cos| + 2 rows of 253 |cos| + 125 |sin| + 2 pixels, i.e. rows x (16-byte chunks per row) x 16 <= 26 688 bytes
// over all rotations (against 40 320 for the whole lattice) -- what lets a fourth CTA live on the SM.
constexpr int SWIN_BYTES = 26688;
constexpr int DWORDS = 5120;            // capacity of the dynamic layer's screen windows, 8 pixels per word
// A trajectory box is drawn for up to 20 steps at the same screen place (only the camera moves), so the run it paints on each of its
// rows is computed once, the step it joins the trail, and kept in HBM beside the pose it belongs to (one record per env and slot of
// the trajectory ring buffer); a later step takes the record iff pose and screen offsets are the ones it would compute from.
// Per env: the 20 slots' runs (short2[64] each, 5 120 B: ONE bulk copy into the span table, requested before anything else),
// then the 20 heads (x, y, cos h, sin h, kbx, kby as float64 + padding).
constexpr int SPANREC_HEAD = 64;
constexpr int SPANREC_RUNS = TRAJ * DROWS * 4;
constexpr int SPANREC = SPANREC_RUNS + TRAJ * SPANREC_HEAD;   // per env
static_assert(SPANREC % 16 == 0 && SPANREC_RUNS % 16 == 0, "bulk-copy source alignment");
constexpr int DHALF = DWORDS / 2;        // the older and the newer half of the boxes are painted into windows of their own

struct SmemDyn {                        // k_render
    Camera cam;
    int fast;                           // every sample of the image maps inside the screen: no per-sample checks
    int ndyn;
    uint32_t probe;                     // palette index of screen pixel (0, 0): rotate()'s background colour
    int wy0[2], nrows[2], cb0[2], nch[2];  // staged window of image half h: screen rows [wy0, wy0 + nrows), 16-byte chunks [cb0, cb0 + nch) of each
    alignas(8) unsigned long long bar;  // mbarrier of the static rows
    alignas(8) unsigned long long bar2; // mbarrier of the span records
    int slot0;                          // ring slot of trajectory box 1 (the oldest drawn)
    alignas(16) Shape shapes[NDYN];     // vehicle box (0), trajectory boxes old -> new (1 ..)
    short2 drange[NDYN];                // (miny, maxy) of dynamic box d
    uchar4 lbox[NDYN];                  // lattice columns [x, y] and rows [z, w] the box can cover (x > y: none)
    uint32_t pal[NCOLOR];               // R | G << 10 | B << 20: four samples add without carry
    uint8_t hit[NDYN + 3];              // the span record of the box's ring slot is valid
    alignas(16) short2 dyn[NDYN][DROWS];  // [ring slot of the trajectory box (see dslot), TRAJ for the vehicle]: span of the box on screen row miny_d + r; x > y: nothing; x == DYN_DIRECT: evaluate
    // dynamic layer as two screen windows (0: boxes [0, dsplit), 1: boxes [dsplit, ndyn), each the bounding box of its boxes cut to
    // the view): rows [dy0, dy0 + dnr), pixels [dx0, dx0 + 8 dnw), dx0 % 8 == 0, words dwin[g * DHALF ..); nibble = 1 + box - first box
    int dx0[2], dy0[2], dnw[2], dnr[2];
    union {
        alignas(16) uint32_t dwin[DWORDS];   // window mode: one nibble per screen pixel and window, the code of the newest box painted there
        alignas(16) uint8_t didx[LAT][LAT];  // lattice mode: palette index of the newest dynamic box on the sample, 0: none
    };
    alignas(16) uint8_t swin[SWIN_BYTES];  // the static screen rows of one image half, 2 bits per pixel, pitch 16 nch
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


}  // namespace render

#ifndef HOPE_RENDER_HOST_TEST  // tests/render_host_harness.cpp compiles the scan-conversion helpers above with g++

// One thread per env: the camera of this step (see Camera), and the env joins `repaint` (list, *repaint_n its length, zeroed before
// the launch) when its scene changed since its static screen was painted (keys: pool slot + the slot's regeneration count).
__global__ void __launch_bounds__(128) k_render_camera(int n, Pool pool, EnvState st, hope_params par, render::Camera *__restrict__ cams,
                                                       const unsigned *__restrict__ episode, const uint2 *__restrict__ keys,
                                                       int *__restrict__ repaint, int *__restrict__ repaint_n) {
    using namespace render;
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    {
        const int sid = st.scene[env];
        const uint2 key = keys[env];
        if (key.x != (unsigned)sid || key.y != (episode ? episode[sid] : 0u)) repaint[atomicAdd(repaint_n, 1)] = env;
    }
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

// A small grid walks the envs k_render_camera listed (none on most steps, all after a reset of the whole batch), one CTA per env
// at a time.  Paints obstacles, start outline and dest box tile by tile (100 screen rows of thermometer bytes
// in shared memory, every (shape, row) pair one work item, spans OR-ed in) and stores the tile as 2 bits per pixel.
__global__ void __launch_bounds__(render::THREADS)
k_render_static(Pool pool, EnvState st, const unsigned *__restrict__ episode, const render::Camera *__restrict__ cams, hope_params par,
                uint8_t *__restrict__ screen, uint2 *__restrict__ keys, const int *__restrict__ repaint, const int *__restrict__ repaint_n) {
    using namespace render;
    extern __shared__ __align__(16) unsigned char render_smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(render_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int nrepaint = *repaint_n;
#pragma unroll 1
    for (int item = blockIdx.x; item < nrepaint; item += gridDim.x) {   // (every pass ends with a block barrier: the tile loop's)
    const int env = repaint[item];
    const int sid = st.scene[env];
    const unsigned ep = episode ? episode[sid] : 0u;
    const double *meta = pool.meta + (size_t)sid * META;
    const int nobs = pool.nobs[sid];
    {   // one thread per shape; start and dest in warps of their own next to the obstacles' (when those fit one warp)
        int s = tid < nobs + 2 ? tid : -1;
        if constexpr (MAXO <= 32) {
            const int w = tid >> 5;
            s = -1;
            if (w == 0) { if (lane < nobs) s = lane; }
            else if (w <= 2 && lane == 0) s = nobs + (w - 1);
        }
        if (tid == 0) sm.nstatic = nobs + 2;
        if (s >= 0) {
            Shape &S = sm.shapes[s];
            double bx[4], by[4];
            Camera cam;  // only the screen offsets are read here
            cam.kbx = cams[env].kbx; cam.kby = cams[env].kby;
            int nv = 4, code, outline = 0;
            if (s < nobs) {  // :303-305 obstacles
                nv = pool.nv[(size_t)sid * MAXO + s];
                const double2 *v = reinterpret_cast<const double2 *>(pool.obs) + ((size_t)sid * MAXO + s) * MAXV;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < nv) { const double2 p = __ldg(v + k); bx[k] = p.x; by[k] = p.y; }
                code = (int)colour_code(1);
            } else if (s == nobs) {  // :307-308 start box, width = 1: lines(closed=True) over the 5 coordinates
                double sn, c;
                sincos(meta[M_START + 2], &sn, &c);
                vehicle_box(meta[M_START], meta[M_START + 1], c, sn, par.box_x, par.box_y, bx, by);
                code = (int)colour_code(2); outline = 1;
            } else {  // :309-310 dest box
#pragma unroll
                for (int k = 0; k < 4; ++k) { bx[k] = meta[M_DBX + k]; by[k] = meta[M_DBY + k]; }
                code = (int)colour_code(3);
            }
            ring_shape(S, cam, bx, by, nv, code, outline);
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
        }
    }
    __syncthreads();
    const int nstatic = sm.nstatic;
    uint32_t *dst = reinterpret_cast<uint32_t *>(screen + (size_t)env * SCREEN_BYTES);
#pragma unroll 1
    for (int t = 0; t < WIN / TROWS; ++t) {
        const int ty0 = t * TROWS, ty1 = ty0 + TROWS - 1;
        for (int k = tid; k < TILEW / 4; k += THREADS) reinterpret_cast<uint4 *>(sm.tile)[k] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        int first = 0;  // rows handed out so far (mod 256): shape after shape the items go round the threads
        for (int sb = 0; sb < nstatic; sb += 32) {  // each lane tests one shape against the tile, the warp walks the hits
            bool touch = false;
            if (sb + lane < nstatic) {
                const Shape &S = sm.shapes[sb + lane];
                touch = S.maxy >= ty0 && S.miny <= ty1 && S.maxx >= 0 && S.minx <= WIN - 1;
            }
            unsigned m = __ballot_sync(HOPE_FULL_MASK, touch);
            while (m) {
                const int s = sb + __ffs(m) - 1;
                m &= m - 1;
                const Shape &S = sm.shapes[s];
                const int lo = max((int)S.miny, ty0), cnt = min((int)S.maxy, ty1) - lo + 1;
                const int tt = (tid - first) & (THREADS - 1);
                first += cnt;
                if (tt < cnt) paint_shape_row(sm, S, lo + tt, sm.tile + (lo + tt - ty0) * TPITCHW, 0, 0, WIN - 1);
            }
        }
        __syncthreads();
        // thermometer bytes (0, 1, 3, 7) -> 2-bit colour indices, 16 pixels per stored word
        for (int k = tid; k < TILEW / 4; k += THREADS) {
            const uint4 w4 = reinterpret_cast<const uint4 *>(sm.tile)[k];
            const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
            uint32_t o = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t l = (ws[q] & 0x01010101u) + ((ws[q] >> 1) & 0x01010101u) + ((ws[q] >> 2) & 0x01010101u);  // per byte: 0..3
                o |= ((l & 3u) | ((l >> 6) & 0xcu) | ((l >> 12) & 0x30u) | ((l >> 18) & 0xc0u)) << (8 * q);
            }
            dst[t * (TILEW / 4) + k] = o;
        }
        __syncthreads();
    }
    if (tid == 0) keys[env] = make_uint2((unsigned)sid, ep);
    }
}

// One CTA per env.
//   1. set-up: the camera thread derives the screen window the 128 x 128 sample lattice can touch; one thread per dynamic box
//      (vehicle + the last <= 20 trajectory boxes) prepares its scan-conversion record and the lattice rectangle it can cover;
//   2. the window's rows of the cached static screen are requested into shared memory (cp.async.bulk, one copy per row, completion
//      counted by one mbarrier) and arrive under steps 3 and 4;
//   3. span table: what each dynamic box paints on each of its screen rows, one (box, row) pair per thread and pass;
//   4. dynamic layer in LATTICE space: newest box first, every thread tests ITS samples (lattice row % 8 = warp, column % 32 = lane: a
//      sample always belongs to the same thread, so program order replaces barriers) of the box's lattice rectangle against the span
//      table and records the box's palette index where nothing newer did;
//   5. gather: an output pixel = 4 samples; a sample = its dynamic index if any, else the 2-bit static pixel of the staged window;
//      palette sums in one packed word, rounded mean, uint8 [3][64][64] with a warp writing whole 32-byte sectors.
// traj: [N][20][4] ring buffer (x, y, cos h, sin h) of Vehicle.trajectory's tail, traj_n: [N] its length (see k_advance);
// traj_len <= TRAJ: how many of them are drawn, in the colours 5 + traj_len - ntraj .. of the palette (TRAJ_COLORS[-ntraj:])
__global__ void __launch_bounds__(render::THREADS, 4)
k_render(int n, EnvState st, const render::Camera *__restrict__ cams, hope_params par, render::Palette pal, const uint8_t *__restrict__ screen,
         uint8_t *__restrict__ img, uint8_t *__restrict__ spanrec, int traj_len, int force_lattice) {
    using namespace render;
    extern __shared__ __align__(16) unsigned char render_smem_raw[];
    SmemDyn &sm = *reinterpret_cast<SmemDyn *>(render_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int env = blockIdx.x;
    const uint8_t *scr = screen + (size_t)env * SCREEN_BYTES;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&sm.bar), bar2 = (unsigned)__cvta_generic_to_shared(&sm.bar2);
    uint8_t *const rec = spanrec + (size_t)env * SPANREC;
    if (tid == 0) {  // the span records of all 20 ring slots, whatever they hold: they are validated when they are here
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar2) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar2), "r"(SPANREC_RUNS) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((unsigned)__cvta_generic_to_shared(&sm.dyn[0][0])), "l"(rec), "r"(SPANREC_RUNS), "r"(bar2) : "memory");
    }
    // ---------------------------------------------------------------- 1. set-up
    static_assert(offsetof(SmemDyn, dwin) % 16 == 0 && offsetof(SmemDyn, swin) % 16 == 0 && DWORDS % (4 * THREADS) == 0 && DWORDS * 4 >= LAT * LAT,
                  "cleared as uint4; bulk copy targets; didx fits");
    static_assert(4 * (sizeof(SmemDyn) + 1024) <= 227 * 1024, "4 CTAs per SM (227 KB, 1 KB reserved per CTA)");
#pragma unroll
    for (int k = 0; k < DWORDS / (4 * THREADS); ++k) reinterpret_cast<uint4 *>(sm.dwin)[tid + k * THREADS] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == THREADS - 2) {
        Camera c = cams[env];
        // crop pixels inside the 500 x 500 blit target AND inside the rotated copy (both axis-aligned in u, v)
        c.ulo = max(-c.rx0, -c.cx0); c.uhi = min(WIN - 1 - c.rx0, c.nx - 1 - c.cx0);
        c.vlo = max(-c.ry0, -c.cy0); c.vhi = min(WIN - 1 - c.ry0, c.ny - 1 - c.cy0);
        sm.cam = c;
        // the sample lattice u, v in {4i+1, 4i+2}: the map is affine, so its corners bound it
        const int fmax_ = (WIN << 16) - 1;
        int fx0 = 0x7fffffff, fx1 = -0x7fffffff - 1, fy0 = 0x7fffffff, fy1 = -0x7fffffff - 1;
        for (int k = 0; k < 4; ++k) {
            const int xc = ((k & 1) ? OBS - 2 : 1) + c.cx0, yc = ((k & 2) ? OBS - 2 : 1) + c.cy0;
            const int fx = c.a0 + c.a1 * xc + c.a2 * yc, fy = c.b0 + c.b1 * xc + c.b2 * yc;
            fx0 = min(fx0, fx); fx1 = max(fx1, fx); fy0 = min(fy0, fy); fy1 = max(fy1, fy);
        }
        sm.fast = (1 >= c.ulo && OBS - 2 <= c.uhi && 1 >= c.vlo && OBS - 2 <= c.vhi && fx0 >= 0 && fy0 >= 0 && fx1 <= fmax_ && fy1 <= fmax_) ? 1 : 0;
        const int wx0 = max(fx0 >> 16, 0), wx1 = min(fx1 >> 16, WIN - 1), wy0 = max(fy0 >> 16, 0), wy1 = min(fy1 >> 16, WIN - 1);
        sm.cam.wx0 = wx0; sm.cam.wy0 = wy0; sm.cam.wx1 = wx1; sm.cam.wy1 = wy1;  // on-screen part of the lattice's bounding box (empty: wx0 > wx1 or wy0 > wy1)
        // staged windows: the on-screen part of the bounding box of either half of the lattice (v = 1 .. 126 and 129 .. 254),
        // whole 16-byte chunks (64 pixels) per row
        for (int h = 0; h < 2; ++h) {
            int gx0 = 0x7fffffff, gx1 = -0x7fffffff - 1, gy0 = 0x7fffffff, gy1 = -0x7fffffff - 1;
            for (int k = 0; k < 4; ++k) {
                const int xc = ((k & 1) ? OBS - 2 : 1) + c.cx0, yc = ((k & 2) ? OBS / 2 - 2 : 1) + h * (OBS / 2) + c.cy0;
                const int fx = c.a0 + c.a1 * xc + c.a2 * yc, fy = c.b0 + c.b1 * xc + c.b2 * yc;
                gx0 = min(gx0, fx); gx1 = max(gx1, fx); gy0 = min(gy0, fy); gy1 = max(gy1, fy);
            }
            const int hx0 = max(gx0 >> 16, 0), hx1 = min(gx1 >> 16, WIN - 1), hy0 = max(gy0 >> 16, 0), hy1 = min(gy1 >> 16, WIN - 1);
            int nrows = 0, nch = 0, cb0 = 0;
            if (hx0 <= hx1 && hy0 <= hy1) {
                cb0 = hx0 >> 6; nch = (hx1 >> 6) - cb0 + 1;
                nrows = min(hy1 - hy0 + 1, SWIN_BYTES / (16 * nch));  // (never the bound: see SWIN_BYTES)
            }
            sm.wy0[h] = hy0; sm.nrows[h] = nrows; sm.cb0[h] = cb0; sm.nch[h] = nch;
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(sm.nrows[0] * sm.nch[0] * 16) : "memory");
    }
    if (tid >= THREADS - 2 - NCOLOR && tid < THREADS - 2) {
        const int k = tid - (THREADS - 2 - NCOLOR);
        sm.pal[k] = (pal.rg[k] & 0xffu) | (((pal.rg[k] >> 16) & 0xffu) << 10) | ((pal.b[k] & 0xffu) << 20);
    }
    {
        const int tn = st.traj_n[env];
        const int ntraj = tn > 1 ? min(tn, traj_len) : 0;  // car_parking_base.py:315-316; traj_len = TRAJ_RENDER_LEN, 0 when RENDER_TRAJ is off
        if (tid == 0) { sm.ndyn = 1 + ntraj; sm.slot0 = (tn - ntraj) % TRAJ; }
        // dynamic box d: 0 = the vehicle (warp 1, lane 0), 1 + i = trajectory box i, old -> new (warp 0)
        int d = -1;
        if (tid < ntraj) d = 1 + tid; else if (tid == 32) d = 0;
        bool hit = false;
        if (d >= 0) {
            Shape &S = sm.shapes[d];
            double bx[4], by[4];
            const Camera cam = cams[env];
            double x, y, c, sn;
            bool skip = false;
            if (d == 0) {  // :312-313 vehicle
                x = st.pose[3 * env]; y = st.pose[3 * env + 1]; c = st.cs[2 * env]; sn = st.cs[2 * env + 1];
                if (ntraj > 0) {  // the newest trajectory box is painted later over the very same pixels: skip this one
                    const double2 *p = reinterpret_cast<const double2 *>(st.traj) + ((size_t)env * TRAJ + (tn - 1) % TRAJ) * 2;
                    const double2 xy = p[0], cs = p[1];
                    skip = xy.x == x && xy.y == y && cs.x == c && cs.y == sn;
                }
            } else {  // :315-319 trajectory[-(ntraj - i)], colour TRAJ_COLORS[-(ntraj - i)]
                const int back = ntraj - (d - 1);  // back = 1: newest
                const double2 *p = reinterpret_cast<const double2 *>(st.traj) + ((size_t)env * TRAJ + (tn - back) % TRAJ) * 2;
                const double2 xy = p[0], cs = p[1];
                x = xy.x; y = xy.y; c = cs.x; sn = cs.y;
                // the span record of this ring slot: valid iff it was computed from this very pose and these screen offsets
                const double2 *h = reinterpret_cast<const double2 *>(rec + SPANREC_RUNS + ((tn - back) % TRAJ) * SPANREC_HEAD);
                const double2 h0 = h[0], h1 = h[1], h2 = h[2];
                hit = h0.x == x && h0.y == y && h1.x == c && h1.y == sn && h2.x == cam.kbx && h2.y == cam.kby;
            }
            sm.hit[d] = hit ? 1 : 0;
            vehicle_box(x, y, c, sn, par.box_x, par.box_y, bx, by);
            ring_shape(S, cam, bx, by, 4, 0, 0);
            if (skip) { S.miny = 1; S.maxy = 0; }
            sm.drange[d] = make_short2(S.miny, S.maxy);
            // lattice rectangle: the box's integer screen corners through the inverse of the camera map, in float.  What the box
            // paints lies within a pixel of their hull (floor / ceil of the crossings) and a sample within a pixel of its screen
            // pixel: 2.5 pixels of margin absorb both and the float rounding.
            uchar4 lb = make_uchar4(1, 0, 1, 0);
            if (!skip) {
                const float a1 = (float)cam.a1, a2 = (float)cam.a2, b1 = (float)cam.b1, b2 = (float)cam.b2;
                const float rdet = 1.0f / (a1 * b2 - a2 * b1);
                float umin = 1e30f, umax = -1e30f, vmin = 1e30f, vmax = -1e30f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int px, py;
                    to_screen(cam, bx[k], by[k], px, py);
                    const float X = ((float)px + 0.5f) * 65536.0f - (float)cam.a0, Y = ((float)py + 0.5f) * 65536.0f - (float)cam.b0;
                    const float u = (b2 * X - a2 * Y) * rdet - (float)cam.cx0, v = (a1 * Y - b1 * X) * rdet - (float)cam.cy0;
                    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                }
                umin -= 2.5f; vmin -= 2.5f; umax += 2.5f; vmax += 2.5f;
                // lattice index a <-> crop pixel 4 (a >> 1) + 1 + (a & 1): [2 floor((umin - 2) / 4), 2 floor(umax / 4) + 1] covers [umin, umax]
                const int iu0 = max(2 * (int)floorf((umin - 2.0f) * 0.25f), 0), iu1 = min(2 * (int)floorf(umax * 0.25f) + 1, LAT - 1);
                const int iv0 = max(2 * (int)floorf((vmin - 2.0f) * 0.25f), 0), iv1 = min(2 * (int)floorf(vmax * 0.25f) + 1, LAT - 1);
                if (iu0 <= iu1 && iv0 <= iv1) lb = make_uchar4((unsigned char)iu0, (unsigned char)iu1, (unsigned char)iv0, (unsigned char)iv1);
            }
            sm.lbox[d] = lb;
        }

    }
    __syncthreads();
    const Camera &cam = sm.cam;
    const int ndyn = sm.ndyn, ntraj = ndyn - 1;
    // ---------------------------------------------------------------- 2. request the static rows of the window
    // every thread that reads what a bulk copy wrote watches the mbarrier phase itself
    auto wait_phase = [&](unsigned b, int parity) {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
    };
    auto stage_rows = [&](int h) {
        const int nrows = sm.nrows[h], bytes = sm.nch[h] * 16;
        const uint8_t *src = scr + (size_t)sm.wy0[h] * SCREEN_PITCH + sm.cb0[h] * 16;
        for (int r = tid; r < nrows; r += THREADS) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.swin[r * bytes]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(src + (size_t)r * SCREEN_PITCH), "r"(bytes), "r"(bar) : "memory");
        }
    };
    stage_rows(0);
    const int dsplit = (ndyn + 1) >> 1;
    static_assert(NDYN <= 30, "(NDYN + 1) / 2 codes per window fit a nibble");
    if (warp < 2) {  // the windows of the dynamic layer, warp g for window g: bounding box of its boxes, cut to what the lattice can touch
        int x0 = 0x7fff, y0 = 0x7fff, x1 = -0x8000, y1 = -0x8000;
        const int d = (warp ? dsplit : 0) + lane;
        if (d < (warp ? ndyn : dsplit)) {
            const Shape &S = sm.shapes[d];
            if (S.miny <= S.maxy) { x0 = S.minx; x1 = S.maxx; y0 = S.miny; y1 = S.maxy; }
        }
        static_assert(NDYN <= 32, "one lane per dynamic box");
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            x0 = min(x0, __shfl_xor_sync(HOPE_FULL_MASK, x0, o)); y0 = min(y0, __shfl_xor_sync(HOPE_FULL_MASK, y0, o));
            x1 = max(x1, __shfl_xor_sync(HOPE_FULL_MASK, x1, o)); y1 = max(y1, __shfl_xor_sync(HOPE_FULL_MASK, y1, o));
        }
        if (lane == 0) {
            x0 = max(x0, cam.wx0); x1 = min(x1, cam.wx1); y0 = max(y0, cam.wy0); y1 = min(y1, cam.wy1);
            int dnw = 0, dnr = 0;
            if (x0 <= x1 && y0 <= y1) { dnw = ((x1 >> 3) - (x0 >> 3) + 1) | 1; dnr = y1 - y0 + 1; }  // odd pitch: the row owners (consecutive rows) hit different banks
            sm.dx0[warp] = x0 & ~7; sm.dy0[warp] = y0; sm.dnw[warp] = dnw; sm.dnr[warp] = dnr;
        }
    }
    // screen pixel (0, 0) of the static screen, for rotate()'s background colour: only read when a sample can leave the screen, and
    // requested here so that it arrives under the span table
    unsigned probe_static = 0u;
    if (tid == THREADS - 1 && !sm.fast) probe_static = scr[0];
    // ---------------------------------------------------------------- 3. span table of the dynamic boxes
    const int slot0 = sm.slot0;
    // span-table row of dynamic box d: its slot of the trajectory ring buffer (the records arrive in ring order), TRAJ for the vehicle
    auto dslot = [&](int d) -> int { const int s = slot0 + d - 1; return d == 0 ? TRAJ : (s >= TRAJ ? s - TRAJ : s); };
    wait_phase(bar2, 0);  // the span records have landed (requested first thing): what is computed below overwrites the stale ones
    for (int idx = tid; idx < ndyn * DROWS; idx += THREADS) {
        const int d = idx / DROWS, r = idx - d * DROWS;
        if (sm.hit[d]) continue;
        const Shape &S = sm.shapes[d];
        const int y = S.miny + r;
        short2 e = make_short2(1, 0);  // nothing on this row
        if (y <= S.maxy && y >= 0 && y < WIN) {
            if (S.miny == S.maxy) e = make_short2(S.minx, S.maxx);
            else {
                // which edges cross the row (row_crossings' test), then the crossings of the usual two only: the first visited
                // rounds down, the second up
                int x0, x1, x2, x3, cnt = 0, ea = 0, eb = 0;
#pragma unroll
                const int last = y == S.maxy ? 1 : 0;  // [ylo, yhi), and the edges that end on the last row count there
                for (int k = 0; k < 4; ++k)
                    if (k < S.ne) {
                        const int t = y - S.e[k].ylo, dy = S.e[k].dy;
                        if ((unsigned)t < (unsigned)dy || (last && t == dy)) { if (cnt == 0) ea = k; else eb = k; ++cnt; }
                    }
                if (cnt == 2) {
                    const Edge A = S.e[ea], B = S.e[eb];
                    x0 = A.xlo + div_round((y - A.ylo) * A.dx, A.dy, A.rdy, false);
                    x1 = B.xlo + div_round((y - B.ylo) * B.dx, B.dy, B.rdy, true);
                    e = make_short2((short)min(x0, x1), (short)max(x0, x1));
                } else if (cnt > 2) {
                    cnt = row_crossings(S, y, x0, x1, x2, x3);
                    if (cnt == 3) {
                        const int lo = min(x0, min(x1, x2)), hi = max(x0, max(x1, x2));
                        e = make_short2((short)lo, (short)(x0 + x1 + x2 - lo - hi));
                    } else if (cnt == 4) e = make_short2(DYN_DIRECT, 0);  // two runs: left to shape_covers
                }
                for (int k = 0; k < S.nh; ++k)  // a horizontal edge on this row (the closing zero-length edge, usually): one more run
                    if (S.hy[k] == y && e.x != DYN_DIRECT) {
                        const int c = min(S.hxa[k], S.hxb[k]), dd = max(S.hxa[k], S.hxb[k]);
                        if (e.x > e.y) e = make_short2((short)c, (short)dd);
                        else if (c <= e.y + 1 && dd >= e.x - 1) e = make_short2((short)min((int)e.x, c), (short)max((int)e.y, dd));  // touches the run: one run
                        else e = make_short2(DYN_DIRECT, 0);
                    }
            }
        }
        const int slot = dslot(d);
        sm.dyn[slot][r] = e;
        if (d > 0) {  // keep it for the steps to come
            reinterpret_cast<short2 *>(rec)[slot * DROWS + r] = e;
            if (r == 0) {
                const double2 *p = reinterpret_cast<const double2 *>(st.traj) + ((size_t)env * TRAJ + slot) * 2;
                double2 *h = reinterpret_cast<double2 *>(rec + SPANREC_RUNS + slot * SPANREC_HEAD);
                h[0] = p[0]; h[1] = p[1]; h[2] = make_double2(cam.kbx, cam.kby);
            }
        }
    }
    __syncthreads();
    const int a0 = cam.a0, a1 = cam.a1, a2 = cam.a2, b0 = cam.b0, b1 = cam.b1, b2 = cam.b2;
    // ---------------------------------------------------------------- 4. dynamic layer
    // a window that does not fit its half of dwin (a trail of more than ~ 150 x 150 pixels per half inside the view): lattice mode
    const bool lattice_mode = sm.dnw[0] * sm.dnr[0] > DHALF || sm.dnw[1] * sm.dnr[1] > DHALF || (force_lattice && sm.dnr[0] + sm.dnr[1] > 0);
    const int dxA = sm.dx0[0], dyA = sm.dy0[0], dwA = lattice_mode ? 0 : sm.dnw[0], drA = lattice_mode ? 0 : sm.dnr[0];
    const int dxB = sm.dx0[1], dyB = sm.dy0[1], dwB = lattice_mode ? 0 : sm.dnw[1], drB = lattice_mode ? 0 : sm.dnr[1];
    if (!lattice_mode) {
        // window mode: a thread owns a window row and paints its window's boxes' runs on it old -> new (the painter's order) as
        // nibbles, whole words between the ends; rows are independent, so there is nothing to wait for.  Lanes 0-15 of warp w own rows
        // 16 w .. 16 w + 15 (+ 128, ...) of window 0, lanes 16-31 the same rows of window 1: the two halves of the boxes go side by
        // side, and a thread walks (ndyn + 1) / 2 boxes at most.  (Measured on B200 before this: one pass per box with a block
        // barrier; rows dealt to warps modulo 8 with four lanes per row; one window for all boxes, whose 21-box chain per row left
        // the other warps waiting at the barrier for a quarter of the kernel.)
        const int g = lane >> 4;
        const int dx0 = g ? dxB : dxA, dy0 = g ? dyB : dyA, dnw = g ? dwB : dwA, dnr = g ? drB : drA;
        const int dfirst = g ? dsplit : 0, dcount = g ? ndyn - dsplit : dsplit;
        uint32_t *win = sm.dwin + g * DHALF;
#pragma unroll 1
        for (int wr = (warp << 4) | (lane & 15); wr < dnr; wr += THREADS / 2) {
            const int y = wr + dy0, xend = 8 * dnw - 1;
            uint32_t *row = win + wr * dnw;
#pragma unroll 1
            for (int k = 0; k < dcount; ++k) {
                const int d = dfirst + k;
                const short2 dr = sm.drange[d];
                if (y < dr.x || y > dr.y) continue;
                const short2 e = sm.dyn[dslot(d)][y - dr.x];
                const uint32_t fill = (uint32_t)(k + 1) * 0x11111111u;
                if (e.x != DYN_DIRECT) {
                    const int na = max((int)e.x - dx0, 0), nb = min((int)e.y - dx0, xend);
                    if (na <= nb) {
                        const int wa = na >> 3, wb = nb >> 3;
                        const uint32_t ma = 0xffffffffu << (4 * (na & 7)), mb = 0xffffffffu >> (4 * (7 - (nb & 7)));
                        if (wa == wb) row[wa] = (row[wa] & ~(ma & mb)) | (fill & ma & mb);
                        else {
                            row[wa] = (row[wa] & ~ma) | (fill & ma);
                            for (int w = wa + 1; w < wb; ++w) row[w] = fill;
                            row[wb] = (row[wb] & ~mb) | (fill & mb);
                        }
                    }
                } else {  // two runs on this row (never for a rectangle): pixel by pixel
                    const Shape &S = sm.shapes[d];
                    for (int x = max((int)S.minx, dx0); x <= min((int)S.maxx, dx0 + xend); ++x)
                        if (shape_covers(S, x, y)) {
                            const int nx = x - dx0;
                            row[nx >> 3] = (row[nx >> 3] & ~(0xfu << (4 * (nx & 7)))) | (fill & (0xfu << (4 * (nx & 7))));
                        }
                }
            }
        }
    } else {
        // lattice mode (a long trail across the whole view): newest box first, every thread tests ITS samples (lattice row % 8 = warp,
        // column % 32 = lane: a sample always belongs to the same thread, so program order replaces barriers) of the box's lattice
        // rectangle against the span table and records the box's palette index where nothing newer did
        const int fx00 = a0 + a1 * cam.cx0 + a2 * cam.cy0, fy00 = b0 + b1 * cam.cx0 + b2 * cam.cy0;
#pragma unroll 1
        for (int d = ndyn - 1; d >= 0; --d) {
            const uchar4 lb = sm.lbox[d];
            if (lb.x > lb.y) continue;
            const int miny = sm.drange[d].x;
            const uint8_t pidx = (uint8_t)(d == 0 ? 4 : 5 + traj_len - ntraj + d - 1);
            const int acol = lb.x + ((lane - lb.x) & 31);
#pragma unroll 1
            for (int b = lb.z + ((warp - lb.z) & 7); b <= lb.w; b += 8) {
                const int v = 2 * b + 1 - (b & 1);
                const int fxr = fx00 + a2 * v, fyr = fy00 + b2 * v;
                for (int a = acol; a <= lb.y; a += 32) {
                    const int u = 2 * a + 1 - (a & 1);
                    const int sx = (fxr + a1 * u) >> 16, sy = (fyr + b1 * u) >> 16;
                    const unsigned r = (unsigned)(sy - miny);
                    if (r < (unsigned)DROWS && sm.didx[b][a] == 0) {
                        const short2 e = sm.dyn[dslot(d)][r];
                        if ((sx >= e.x && sx <= e.y) || (e.x == DYN_DIRECT && shape_covers(sm.shapes[d], sx, sy))) sm.didx[b][a] = pidx;
                    }
                }
            }
        }
    }
    if (tid == THREADS - 1 && !sm.fast) {  // screen pixel (0, 0), rotate()'s background colour
        unsigned idx = probe_static & 3u;
        for (int d = 0; d < ndyn; ++d)
            if (sm.shapes[d].miny <= 0 && sm.shapes[d].minx <= 0 && shape_covers(sm.shapes[d], 0, 0)) idx = d == 0 ? 4u : (unsigned)(5 + traj_len - ntraj + d - 1);
        sm.probe = idx;
    }
    __syncthreads();
    wait_phase(bar, 0);  // the static rows of the first half have landed
    // ---------------------------------------------------------------- 5. gather 64 x 64 x (2 x 2 samples), one half of the image at a time
    {
        // a warp resolves a tile of 8 x 4 output pixels per round (16 rounds): a compact patch of the screen, so that few warps
        // meet the dynamic layer's window; its stores are 4 x 8 bytes per channel and merge with the neighbour tiles' in L2
        const int il = lane & 7, jl = lane >> 3;
        uint8_t *out = img + (size_t)env * 3 * IMG * IMG;
        const uint32_t round2 = 2u | (2u << 10) | (2u << 20);  // (a + b + c + d + 2) >> 2 per 10-bit lane
        // palette indices of the dynamic layer on the four samples of output pixel (i, j), one per byte, 0: none
        auto dyn_of = [&](int i_, int j_, int fxb, int fyb) -> unsigned {
            if (lattice_mode)
                return (unsigned)*reinterpret_cast<const unsigned short *>(&sm.didx[2 * j_][2 * i_]) |
                       ((unsigned)*reinterpret_cast<const unsigned short *>(&sm.didx[2 * j_ + 1][2 * i_]) << 16);
            // the four samples lie within two pixels of the first
            const int sx0 = fxb >> 16, sy0 = fyb >> 16;
            const bool nearA = (unsigned)(sy0 - dyA + 2) < (unsigned)(drA + 4) && (unsigned)(sx0 - dxA + 2) < (unsigned)(8 * dwA + 4);
            const bool nearB = (unsigned)(sy0 - dyB + 2) < (unsigned)(drB + 4) && (unsigned)(sx0 - dxB + 2) < (unsigned)(8 * dwB + 4);
            if (!nearA && !nearB) return 0u;
            unsigned out4 = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int fx = fxb + ((k & 1) ? a1 : 0) + ((k >> 1) ? a2 : 0), fy = fyb + ((k & 1) ? b1 : 0) + ((k >> 1) ? b2 : 0);
                const int sx = fx >> 16, sy = fy >> 16;
                int d = -1;  // the newer half's window first
                {
                    const int rx = sx - dxB, ry = sy - dyB;
                    if ((unsigned)ry < (unsigned)drB && (unsigned)rx < (unsigned)(8 * dwB)) {
                        const unsigned c = (sm.dwin[DHALF + ry * dwB + (rx >> 3)] >> (4 * (rx & 7))) & 15u;
                        if (c) d = dsplit + (int)c - 1;
                    }
                }
                if (d < 0) {
                    const int rx = sx - dxA, ry = sy - dyA;
                    if ((unsigned)ry < (unsigned)drA && (unsigned)rx < (unsigned)(8 * dwA)) {
                        const unsigned c = (sm.dwin[ry * dwA + (rx >> 3)] >> (4 * (rx & 7))) & 15u;
                        if (c) d = (int)c - 1;
                    }
                }
                if (d >= 0) out4 |= (d == 0 ? 4u : (unsigned)(5 + traj_len - ntraj + d - 1)) << (8 * k);
            }
            return out4;
        };
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
        if (half) {  // the other half's rows through the same buffer, once everybody is done with the first
            // The arrival that opens phase 1 comes AFTER the block barrier: every thread has left its phase-0 wait by then (an empty
            // second window completes phase 1 at once, and a thread still polling parity 0 would then wait for phase 2 for ever).
            // Copies of other threads may complete before it is posted: the transaction count may go negative in between.
            __syncthreads();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tid == THREADS - 2) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(sm.nrows[1] * sm.nch[1] * 16) : "memory");
            stage_rows(1);
            wait_phase(bar, 1);
        }
        const int wpitch = sm.nch[half] * 16;
        const int woff = sm.wy0[half] * wpitch + sm.cb0[half] * 16;  // static pixel (sx, sy): 2 bits of swin[sy * wpitch + (sx >> 2) - woff]
        const int r0 = half * (IMG * IMG / THREADS / 2), r1 = r0 + IMG * IMG / THREADS / 2;
        if (sm.fast) {
#pragma unroll 2
            for (int r = r0; r < r1; ++r) {
                const int i = ((warp & 7) << 3) | il, j = (r << 2) | jl;
                const int xc = 4 * i + 1 + cam.cx0, yc = 4 * j + 1 + cam.cy0;
                const int fxb = a0 + a1 * xc + a2 * yc, fyb = b0 + b1 * xc + b2 * yc;
                const unsigned dyn4 = dyn_of(i, j, fxb, fyb);
                unsigned st4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int fx = fxb + ((k & 1) ? a1 : 0) + ((k >> 1) ? a2 : 0), fy = fyb + ((k & 1) ? b1 : 0) + ((k >> 1) ? b2 : 0);
                    const int sx = fx >> 16, sy = fy >> 16;
                    st4[k] = ((unsigned)sm.swin[sy * wpitch + (sx >> 2) - woff] >> (2 * (sx & 3))) & 3u;
                }
                uint32_t sum = round2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned dk = (dyn4 >> (8 * k)) & 0xffu;
                    sum += sm.pal[dk ? dk : st4[k]];
                }
                uint8_t *o = out + j * IMG + i;
                o[0] = (uint8_t)((sum >> 2) & 0xffu); o[IMG * IMG] = (uint8_t)((sum >> 12) & 0xffu); o[2 * IMG * IMG] = (uint8_t)((sum >> 22) & 0xffu);
            }
        } else {
            const unsigned bgidx = sm.probe;
            const unsigned xmaxv = (WIN << 16) - 1;
            const int ulo = cam.ulo, uhi = cam.uhi, vlo = cam.vlo, vhi = cam.vhi;
#pragma unroll 1
            for (int r = r0; r < r1; ++r) {
                const int i = ((warp & 7) << 3) | il, j = (r << 2) | jl;
                const int ub = 4 * i + 1, xc = ub + cam.cx0;
                const int vb = 4 * j + 1, yc = vb + cam.cy0;
                const int fxb = a0 + a1 * xc + a2 * yc, fyb = b0 + b1 * xc + b2 * yc;
                const unsigned dyn4 = dyn_of(i, j, fxb, fyb);
                uint32_t sum = round2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int u = ub + (k & 1), v = vb + (k >> 1);
                    unsigned idx = 0u;  // background: white -> black, and the untouched (black) blit target
                    if (u >= ulo && u <= uhi && v >= vlo && v <= vhi) {
                        const int fx = fxb + ((k & 1) ? a1 : 0) + ((k >> 1) ? a2 : 0), fy = fyb + ((k & 1) ? b1 : 0) + ((k >> 1) ? b2 : 0);
                        if ((unsigned)fx > xmaxv || (unsigned)fy > xmaxv) idx = bgidx;  // negative wraps to a huge unsigned
                        else {
                            idx = (dyn4 >> (8 * k)) & 0xffu;
                            if (idx == 0u) {
                                const int sx = fx >> 16, sy = fy >> 16;
                                idx = ((unsigned)sm.swin[sy * wpitch + (sx >> 2) - woff] >> (2 * (sx & 3))) & 3u;
                            }
                        }
                    }
                    sum += sm.pal[idx];
                }
                uint8_t *o = out + j * IMG + i;
                o[0] = (uint8_t)((sum >> 2) & 0xffu); o[IMG * IMG] = (uint8_t)((sum >> 12) & 0xffu); o[2 * IMG * IMG] = (uint8_t)((sum >> 22) & 0xffu);
            }
        }
        }  // half
    }
}

#endif  // HOPE_RENDER_HOST_TEST
