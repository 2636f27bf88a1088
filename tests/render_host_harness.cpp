// CPU harness for tests/test_render_host.py: compiles the scan-conversion helpers of hope_b200/csrc/render.cuh
// (ring_shape, make_seg, row_crossings, div_round, span, paint_shape_row, shape_covers — the code k_render runs per
// row) with g++, so they can be fuzzed against the raster oracle over far more shapes than the GPU tests visit
// (triangles, slivers, horizontal edges, concave and self-intersecting quads, outlines).  The CUDA intrinsics they use
// are given their documented host meaning here.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define HOPE_RENDER_HOST_TEST 1
#define __device__
#define __host__
#define __noinline__
#define __forceinline__ inline
struct short2 { short x, y; };
struct uint2 { unsigned x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline short2 make_short2(short x, short y) { return short2{x, y}; }
template <class A, class B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <class A, class B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { const uint32_t old = *p; *p = old | v; return old; }  // one host thread
static inline float __frcp_rn(float x) { return 1.0f / x; }             // IEEE round-to-nearest reciprocal
static inline float __int2float_rn(int x) { return (float)x; }
static inline int __float2int_rd(float x) { return (int)std::floor(x); }
constexpr int MAXO = 16;  // HOPE_MAX_OBS of the default build

#include "../hope_b200/csrc/render.cuh"

using namespace render;

// Paint one ring (nv = 3 or 4 integer screen vertices; closed like shapely's coordinate list) on a 500 x 500 byte
// image with colour index `color`, the way k_render's row owners do: every row through paint_shape_row.
extern "C" int render_host_paint(const int *px, const int *py, int nv, int color, int outline, unsigned char *img) {
    static Smem sm;
    Camera cam;
    std::memset(&cam, 0, sizeof(cam));
    cam.kbx = 0.0; cam.kby = 0.0;
    double bx[4], by[4];
    for (int k = 0; k < nv; ++k) {  // world coordinates that truncate to exactly (px, py) under to_screen
        bx[k] = (px[k] + (px[k] >= 0 ? 0.5 : -0.5)) / KSCALE;
        by[k] = (py[k] + (py[k] >= 0 ? 0.5 : -0.5)) / KSCALE;
        int ix, iy;
        to_screen(cam, bx[k], by[k], ix, iy);
        if (ix != px[k] || iy != py[k]) return -1;
    }
    Shape &S = sm.shapes[0];
    ring_shape(S, cam, bx, by, nv, color, outline);
    if (outline) {
        if (nv != 4) return -2;
        for (int k = 0; k < 4; ++k) make_seg(sm.seg[k], px[k], py[k], px[(k + 1) & 3], py[(k + 1) & 3]);
        make_seg(sm.seg[4], px[0], py[0], px[0], py[0]);
    }
    static uint32_t row[WIN / 4];
    for (int y = 0; y < WIN; ++y) {
        std::memcpy(row, img + (size_t)y * WIN, WIN);
        paint_shape_row(sm, S, y, row, 0, 0, WIN - 1);
        std::memcpy(img + (size_t)y * WIN, row, WIN);
    }
    return 0;
}

// shape_covers for every pixel (the exact point query used for the background probe and two-run rows)
extern "C" int render_host_covers(const int *px, const int *py, int nv, unsigned char *img) {
    Camera cam;
    std::memset(&cam, 0, sizeof(cam));
    double bx[4], by[4];
    for (int k = 0; k < nv; ++k) {
        bx[k] = (px[k] + (px[k] >= 0 ? 0.5 : -0.5)) / KSCALE;
        by[k] = (py[k] + (py[k] >= 0 ? 0.5 : -0.5)) / KSCALE;
    }
    Shape S;
    ring_shape(S, cam, bx, by, nv, 1, 0);
    for (int y = 0; y < WIN; ++y)
        for (int x = 0; x < WIN; ++x) img[(size_t)y * WIN + x] = shape_covers(S, x, y) ? 1 : 0;
    return 0;
}
