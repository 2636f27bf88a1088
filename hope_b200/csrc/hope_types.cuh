// hope_types.cuh — compile-time sizes and the plain argument structs of the kernels (pointers into the SoA arrays of
// DESIGN.md §3).  A header of its own so the host harnesses under tests/ see the same declarations; included by
// hope_kernels.cu inside namespace hope.
#pragma once

constexpr int MAXO = HOPE_MAX_OBS;
constexpr int MAXV = HOPE_MAX_VERTS;
constexpr int MAXE = MAXO * MAXV;  // 64 edges (512 in the obs128 build)
static_assert(MAXO <= 256 && MAXE % 32 == 0, "obstacle indices are packed into 8 bits, edges staged 32 at a time");
constexpr int NRAY = HOPE_N_LIDAR;
constexpr int NACT = HOPE_N_ACTION;
constexpr int NITER = HOPE_N_MASK_ITER;
constexpr int NUP = HOPE_N_UPSAMPLE;
constexpr int META = 24;       // doubles of per-scene metadata
constexpr int MAXW = 16;       // admitted Reeds-Shepp words kept per env
constexpr int ADV_THREADS = 64;

// per-scene metadata layout (doubles)
enum { M_START = 0, M_DEST = 3, M_BOUNDS = 6, M_DBX = 10, M_DBY = 14, M_DAREA = 18, M_DNORM = 19, M_DAABB = 20 };

struct Pool {
    const double *obs;    // [P][16][4][2]
    const uint8_t *nv;    // [P][16]
    const double *aabb;   // [P][16][4] xmin xmax ymin ymax
    const double *meta;   // [P][24]
    const int *nobs;      // [P]
    int size;
};
struct Tables {
    const double *ray_a, *ray_b, *lidar_base, *mask_base;
    const double *dist_star;  // [1200][42][10]
    const double *pmaxk;      // [1200][10][42] running max over k of dist_star, action index contiguous
    const double *pmax;       // [1200]       max_{j,k} dist_star
    const double *gpmax;      // [120]        max of pmax over the 10 upsampled rays of a beam
    const double *w_lo, *w_hi;
    double maxc;
};
struct EnvState {
    double *pose;     // [N][3]
    double *cs;       // [N][2] cos, sin of the heading (k_advance -> k_observe)
    int *t;           // [N]
    double *accum;    // [N]
    int *scene;       // [N]
    uint8_t *pending; // [N] finished last step, takes its next scene on this one
    uint8_t *gate;    // [N] RS gate of this step
    unsigned long long *counters;  // [8]
    double *traj;     // [N][20][4] ring buffer (x, y, cos h, sin h): tail of Vehicle.trajectory (vehicle.py:121-157), read by k_render
    int *traj_n;      // [N] len(Vehicle.trajectory); entry j of the list lives in slot j % 20
};
struct RsWord;   // rs_words.cuh
struct WordSlot;
struct RsScratch {
    RsWord *words;       // [N][MAXW] in try order
    uint8_t *ntry;       // [N]
    uint8_t *ncand;      // [N]
    int *item_base;      // [N]   first work item of env i (its ntry items are consecutive, in try order)
    int *items;          // [N*MAXW] work item -> (env << 4) | try slot
    uint8_t *item_bad;   // [N*MAXW] 1 = the word leaves the map or touches an obstacle
    WordSlot *slots;     // [N*MAXW] per-item sampling plan written by k_rs_walk
    int *n_items;        // [1]   items of this step (reset by the host before k_rs_enumerate)
};
