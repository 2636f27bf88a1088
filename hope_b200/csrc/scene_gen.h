// scene_gen.h — device scene generator entry used by hope_kernels.cu (definition in scene_gen.cu)
#pragma once
#include <stdint.h>

#include "../../include/hope_b200.h"

namespace hope_scene {
// Fill pool slots with freshly generated scenes on the device.  slots == nullptr: slots first..first+n-1;
// otherwise slots[t] (negative = skip).  level_or_mix: 0/1/2, or -1 for slot % 3.  episode (optional,
// [pool]) is mixed into the stream index so a slot regenerates a different scene each time.
int launch_generate(int n, int first, int level_or_mix, uint64_t seed, const int *d_slots, const unsigned *d_episode, const hope_params &par,
                    double *obs, uint8_t *nv, double *aabb, double *meta, int *nobs, int *d_status, void *stream);
}
