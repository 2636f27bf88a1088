// host_wire.h — host-side expansion of hope_step_host's narrow wire format (see host_wire.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace hope_wire {
// float64 action mask rows [lo, hi) from their uint8 step counts (42 per env)
void expand_mask(const uint8_t *steps, double *mask, size_t lo, size_t hi, int force_portable);
// float64 lidar rows [lo, hi): bits[i][4] flags the beams whose value travelled, packed + off[i] is env i's first kept value
// (offsets are relative to `packed`), every other beam reads nohit[ray]
void expand_lidar(const uint32_t *bits, const uint32_t *off, const double *packed, const double *nohit, double *lidar, size_t lo, size_t hi, int force_portable);
int vector_path();  // 1 = the AVX-512 routines are in use on this CPU
}  // namespace hope_wire
