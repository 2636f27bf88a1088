// observe.cuh — shared-memory scratch of k_observe (one per warp = per env).  The statements each warp runs are in
// observe_body.inc, included textually into the kernel (hope_kernels.cu) and into tests/observe_host_harness.cpp, which
// compiles them with g++ on the CPU warp emulation (tests/warp_emu.h) and replays lidar / mask traces recorded from the
// unmodified reference.  (A textual fragment rather than a function: the kernel's machine code stays byte-identical to
// the build that was verified on the GPU.)  Needs hope_types.cuh, hope_device.cuh; included inside namespace hope.
#pragma once

struct ObserveSmem {
    // per edge, packed for 128-bit broadcast loads in the ray loop: line coefficients d x + e y + f = 0 (ego frame)
    // and the edge bounding box
    double2 de[MAXE];      // d, e
    double2 fxn[MAXE];     // f, xmin
    double2 xym[MAXE];     // xmax, ymin
    double eymax[MAXE];
    double L[NRAY];                                              // clip(lidar)+mask_base
    int steps[NACT + 2];
    uint8_t quad[MAXE];                                          // which ray quadrants can accept this edge
    uint16_t qlist[4][MAXE];                                     // per quadrant: the edges that can be hit from it
    uint16_t qcount[4];
};

#include "div_pair.cuh"
