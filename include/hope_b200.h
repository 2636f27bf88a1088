/* hope_b200.h — C ABI of the B200-native batched ParkingEnv step.
 *
 * The reference (jiamiya/HOPE) has no FFI: its boundary is the Python class surface
 *   CarParkingWrapper.reset / .step          src/env/env_wrapper.py:58-85
 *   CarParking.reset / .step                 src/env/car_parking_base.py:127-138, 235-299
 * This header is what a binding for that path would call (see INTEGRATION.md for the ctypes
 * stub).  Plain pointers and sizes only; no torch / CUDA types (a stream is passed as void*).
 *
 * Conventions
 *   - every function returns HOPE_OK (0) or a negative hope_status; nothing throws;
 *   - "d_" parameters are DEVICE pointers owned by the caller (e.g. torch tensors),
 *     "h_" parameters are HOST pointers; all floating point data is float64 like the reference;
 *   - step functions are asynchronous on `stream` (a cudaStream_t, NULL = default stream);
 *   - one context per GPU; a context is not thread-safe.
 *
 * Scene capacity: HOPE_MAX_OBS obstacle rings of up to HOPE_MAX_VERTS vertices.
 */
#ifndef HOPE_B200_H
#define HOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef HOPE_MAX_OBS
#define HOPE_MAX_OBS 16      /* obstacle rings per scene.  The library is also built with -DHOPE_MAX_OBS=128
                                (libhope_b200_obs128.so) for the Dragon Lake Parking scenes (37-125 rings, row f3);
                                hope_max_obs() reports the capacity of the loaded build */
#endif
#define HOPE_MAX_VERTS 4
#define HOPE_N_LIDAR 120     /* configs.py:96  LIDAR_NUM */
#define HOPE_N_ACTION 42     /* configs.py:115 N_DISCRETE_ACTION */
#define HOPE_N_MASK_ITER 10  /* action_mask.py:9 n_iter */
#define HOPE_N_UPSAMPLE 1200 /* LIDAR_NUM * up_sample_rate, action_mask.py:18 */
#define HOPE_RS_MAX_SEG 5

typedef enum {
    HOPE_OK = 0,
    HOPE_ERR_INVALID = -1,     /* bad argument */
    HOPE_ERR_CUDA = -2,        /* a CUDA runtime call failed (hope_last_cuda_error) */
    HOPE_ERR_NO_TABLES = -3,   /* hope_upload_tables not called yet */
    HOPE_ERR_NO_SCENES = -4,   /* hope_set_scene_pool / hope_reset not called yet */
    HOPE_ERR_CAPACITY = -5     /* scene exceeds HOPE_MAX_OBS / HOPE_MAX_VERTS, or the vehicle box is too large for the
                                  image stage's per-box span table */
} hope_status;

/* vehicle.py:13-18 */
enum { HOPE_CONTINUE = 1, HOPE_ARRIVED = 2, HOPE_COLLIDED = 3, HOPE_OUTBOUND = 4, HOPE_OUTTIME = 5 };
/* Reeds-Shepp segment codes in rs_types (reeds_shepp.py ctypes 'S','L','R'); 255 = unused slot */
enum { HOPE_RS_S = 0, HOPE_RS_L = 1, HOPE_RS_R = 2, HOPE_RS_NONE = 255 };

/* stages of one env step (bit mask) */
enum {
    HOPE_STAGE_ADVANCE = 1,  /* kinematics + collision + arrival + status + reward (always runs) */
    HOPE_STAGE_OBSERVE = 2,  /* LiDAR raycast + action-mask sweep + target representation */
    HOPE_STAGE_RS = 4,       /* Reeds-Shepp search (car_parking_base.py:293-297 gate) */
    HOPE_STAGE_ALL = 7,      /* the lidar + action-mask + Reeds-Shepp step (use_img_observation=False) */
    HOPE_STAGE_IMAGE = 8,    /* ego-centric image observation into hope_out.img (car_parking_base.py:301-350,
                                observation_processor.py:6-23); needs a non-NULL img pointer */
    HOPE_STAGE_RAW_ACTION = 16 /* modifier: the action is CarParking.step's [steer rad, speed m/s] (car_parking_base.py:235),
                                not the wrapper's policy output in [-1,1]^2: skip the rescale of env_wrapper.py:37-50 */
};
#define HOPE_IMG_C 3         /* observation_processor.py:9  n_channels */
#define HOPE_IMG_HW 64       /* configs.py:89-90 OBS_W, OBS_H = 256, observation_processor.py:8 downsample_rate 4 */
#define HOPE_N_COLOR 25      /* palette entries: background, obstacle, start, dest, vehicle, 20 trajectory colours */

/* Mirrors the constants of src/configs.py the path reads. */
typedef struct hope_params {
    double wheel_base;        /* configs.py:13  WHEEL_BASE 2.8 */
    double box_x[4], box_y[4];/* configs.py:20-24 VehicleBox corners rb, rf, lf, lb */
    double valid_speed[2];    /* configs.py:32 */
    double valid_steer[2];    /* configs.py:33 */
    int num_step;             /* configs.py:37  NUM_STEP 10 */
    double step_length;       /* configs.py:38  STEP_LENGTH 0.05 */
    int mini_iter;            /* vehicle.py:66  20 */
    double lidar_range;       /* configs.py:95  10 */
    int tolerant_time;        /* configs.py:98  200 */
    double rs_max_dist;       /* configs.py:104 10 */
    double rs_step;           /* car_parking_base.py:424 sampling interval 0.1 m */
    double reward_weight[5];  /* configs.py:181-187 time, rs_dist, dist, angle, box_union */
    double reward_ratio;      /* configs.py:180 0.1 */
    int env_collide;          /* configs.py:79  ENV_COLLIDE: 1 = a collision on the first substep ends the episode
                                 (status COLLIDED, car_parking_base.py:264-267, 279-282); default 0 */
    int auto_reset;           /* 1: an env that finished takes its next pool scene on the following step */
    int regen_on_reset;       /* 1 (needs pool_size >= n_envs): env i owns pool slot i and, when it finishes, a fresh
                                 scene is generated ON THE DEVICE into that slot (scope row f4) instead of cycling
                                 through a pre-generated pool */
    int regen_level;          /* level of regenerated scenes: 0 Normal, 1 Complex, 2 Extrem, -1 = slot % 3 */
    uint64_t regen_seed;      /* stream seed of regenerated scenes */
} hope_params;

/* Per-step outputs; DEVICE pointers, row-major [n_envs][...].  A NULL member is skipped. */
typedef struct hope_out {
    double *pose;         /* [n][3] x, y, heading after the step */
    double *lidar;        /* [n][120]  lidar_simulator.py:31-46 (range minus own-box offset) */
    double *mask;         /* [n][42]   action_mask.py:166-184 (steps/10, or all 0.01) */
    uint8_t *mask_steps;  /* [n][42]   integer collision-free steps 0..10 after the min filter */
    double *target;       /* [n][5]    car_parking_base.py:372-381 */
    double *reward;       /* [n]       env_wrapper.py:10-35 shaped scalar */
    double *reward_info;  /* [n][5]    car_parking_base.py:285-289 */
    int32_t *status;      /* [n]       HOPE_CONTINUE.. */
    uint8_t *done;        /* [n]       status != CONTINUE */
    uint8_t *substeps;    /* [n]       Vehicle.step calls made (0..10) */
    uint8_t *retreated;   /* [n]       1 if the loop ended in a collision retreat */
    uint8_t *was_reset;   /* [n]       1 if this step was an auto-reset (no motion, t = 1) */
    uint8_t *rs_found;    /* [n]       info['path_to_dest'] is not None */
    uint8_t *rs_nseg;     /* [n] */
    uint8_t *rs_types;    /* [n][5] */
    double *rs_lengths;   /* [n][5]    signed metres (PATH.lengths) */
    double *rs_L;         /* [n]       PATH.L */
    uint8_t *rs_ncand;    /* [n]       admissible words (diagnostic) */
    uint8_t *rs_ntried;   /* [n]       words sampled and checked (diagnostic) */
    uint8_t *img;         /* [n][3][64][64]  image observation after Obs_Processor.process_img and the wrapper's
                                       HWC -> CHW transpose, as the uint8 value cv2.resize produced: the reference's
                                       float64 obs['img'] is exactly img / 255.0 (env_wrapper.py:52-55) */
} hope_out;

/* Host mirror of hope_out for hope_step_host (same shapes, HOST pointers, NULL = skip). */
typedef hope_out hope_host_out;

typedef struct hope_ctx hope_ctx;

int hope_default_params(hope_params *p);
int hope_create(hope_ctx **out, int device, int n_envs, int pool_size, const hope_params *p);
int hope_destroy(hope_ctx *ctx);
const char *hope_strerror(int status);
const char *hope_last_cuda_error(const hope_ctx *ctx);

/* Scene-independent tables, HOST pointers (built by the host layer with the reference's numpy
 * expressions; model/action_mask.py:9-163, env/lidar_simulator.py:14-53, 86-88):
 *   ray_a[120]=sin(theta_i)  ray_b[120]=-cos(theta_i)  lidar_base[120]  mask_base[120]
 *   dist_star[1200][42][10]  w_lo[10]=1-r/10  w_hi[10]=r/10 */
int hope_upload_tables(hope_ctx *ctx, const double *h_ray_a, const double *h_ray_b, const double *h_lidar_base,
                       const double *h_mask_base, const double *h_dist_star, const double *h_w_lo,
                       const double *h_w_hi);

/* Colours of the image observation (configs.py:26-30, 80-88), HOST pointer rgb[25][3] in painter's order:
 * [0] BG_COLOR [1] OBSTACLE_COLOR [2] START_COLOR [3] DEST_COLOR [4] vehicle colour COLOR_POOL[0]
 * [5..24] TRAJ_COLORS[0..19].  A palette entry equal to BG_COLOR reads as black in the observation
 * (Obs_Processor.change_bg_color).  hope_create installs the reference's defaults. */
int hope_set_palette(hope_ctx *ctx, const uint8_t *h_rgb);
/* How many of the last trajectory boxes _render draws: configs.py:86 TRAJ_RENDER_LEN (0..20, default 20), 0 for configs.py:105
   RENDER_TRAJ = False (car_parking_base.py:315-320).  Box i of the n drawn (old -> new) takes palette entry 5 + len - n + i,
   i.e. the palette's trajectory colours are TRAJ_COLORS of that length in entries 5 .. 5 + len - 1. */
int hope_set_render_traj(hope_ctx *ctx, int traj_render_len);

/* Scene pool, HOST pointers: scenes [first, first+n) of the pool.
 *   start[n][3] dest[n][3] bounds[n][4]=(xmin,xmax,ymin,ymax) obs_xy[n][16][4][2] nverts[n][16]
 * (ParkingMapNormal fields, parking_map_normal.py:460-494). */
int hope_set_scene_pool(hope_ctx *ctx, int first, int n, const double *h_start, const double *h_dest,
                        const double *h_bounds, const double *h_obs_xy, const int32_t *h_nverts);

/* Procedural scenes on the host (bay / parallel cases, parking_map_normal.py:40-457), own RNG
 * stream per scene (seed + index): level 0 Normal, 1 Complex, 2 Extrem.  Fills HOST arrays. */
int hope_generate_scenes(int n, int level, uint64_t seed, int nthreads, double *h_start, double *h_dest,
                         double *h_bounds, double *h_obs_xy, int32_t *h_nverts, int32_t *h_case_id);

/* Same generator, run on the device (one thread per scene) straight into pool slots [first, first+n);
 * level -1 cycles the three levels by slot % 3.  Asynchronous on `stream`. */
int hope_generate_scene_pool_device(hope_ctx *ctx, int first, int n, int level, uint64_t seed, void *stream);
/* Read pool scenes back to HOST arrays (shapes as in hope_set_scene_pool; synchronous). */
int hope_get_scene_pool(hope_ctx *ctx, int first, int n, double *h_start, double *h_dest, double *h_bounds, double *h_obs_xy,
                        int32_t *h_nverts);

/* env.reset for every env: env i takes pool scene h_scene_ids[i] (NULL: scene i % pool),
 * pose = start, t = 0, accum = 0, then the reset step (no action, t becomes 1; RS skipped). */
int hope_reset(hope_ctx *ctx, const int32_t *h_scene_ids, const hope_out *d_out, void *stream);

/* One CarParkingWrapper.step for every env.  d_action[n][2]: policy output in [-1,1]^2
 * (env_wrapper.py:37-50 rescale happens on the device), or with HOPE_STAGE_RAW_ACTION the physical [steer, speed] of
 * CarParking.step.  d_action == NULL is CarParking.step(None) (car_parking_base.py:255): no motion, t += 1, the full
 * observation, status, reward and (when t > 1) the Reeds-Shepp search from the current pose. */
int hope_step(hope_ctx *ctx, const double *d_action, const hope_out *d_out, unsigned stages, void *stream);

/* BASELINE cfg 2: kinematics + collision (+ arrival) only. */
int hope_step_kinematics_collision(hope_ctx *ctx, const double *d_action, double *d_pose, uint8_t *d_collided,
                                   uint8_t *d_substeps, void *stream);

/* Same step with HOST buffers: copies h_action in, runs, copies the non-NULL outputs back,
 * synchronises.  This is the end-to-end entry point a host-only caller (the reference's
 * training loop) would use.  The host calls order themselves behind the last device-API call on this context
 * (hope_reset / hope_step / hope_planner_actions on the caller's stream), so the two APIs can be mixed. */
int hope_step_host(hope_ctx *ctx, const double *h_action, const hope_host_out *h_out, unsigned stages);
int hope_reset_host(hope_ctx *ctx, const int32_t *h_scene_ids, const hope_host_out *h_out);

/* Host-side half of the narrow wire format of hope_step_host: the float64 action mask of n envs from their uint8
 * step counts, exactly as action_mask.py:182-183 / k_observe compute it (steps / 10, or 0.01 for all 42 actions of an
 * env whose counts are all 0).  HOST pointers h_steps[n][42], h_mask[n][42]; pure host code, no CUDA call. */
int hope_expand_mask(const uint8_t *h_steps, double *h_mask, int n);
int hope_expand_mask_portable(const uint8_t *h_steps, double *h_mask, int n); /* the same without the AVX-512 routine (tests) */

/* The other narrow array of hope_step_host: the float64 lidar (lidar_simulator.py:31-135; 960 of the 1 436 bytes a step returns
 * per env).  A beam that hits nothing within range reads exactly lidar_range - lidar_base[ray] (:46, :134), so on the device
 * k_pack_lidar keeps only the values whose bits differ from that constant and the host rebuilds the rows, bit for bit:
 * h_bits[n][4] (bit j of word q = beam 32 q + j travelled), h_off[n] (index of env i's first kept value in h_packed),
 * h_nohit[120] (the per-ray constant), h_lidar[n][120] (result).  h_packed must stay readable for 64 bytes behind its last
 * kept value (the vector routine loads 8 values at a time).  `portable` != 0 skips the AVX-512 routine.  Pure host code. */
int hope_expand_lidar(const uint32_t *h_bits, const uint32_t *h_off, const double *h_packed, const double *h_nohit, double *h_lidar,
                      int n, int portable);

/* What the last hope_step_host moved and how: info[0] = host-to-device bytes, [1] = device-to-host bytes (all copies of the
 * step, including the data-dependent kept lidar values), [2] / [3] = 1 when the mask / lidar travelled narrow, [4] = host
 * expansion threads, [5] = 1 when the AVX-512 expansion routines are in use, [6] = env ranges the step was pipelined over, [7] = envs whose lidar
 * rows travelled packed (the others' float64 rows were copied as they are: HOPE_B200_HOST_PACK_FRAC, default by ranks per box). */
int hope_host_wire_info(const hope_ctx *ctx, uint64_t info[8]);

/* Batched RsPlanner + ParkingAgent hand-off (model/agent/parking_agent.py:2-47, 64-70, 93-110;
 * train_HOPE_sac.py:194-213).  Call once per rollout step BEFORE hope_step, with the outputs of the
 * previous step: an env whose last step ended (done / was_reset) drops its plan; an env without a plan
 * whose last step found a path (rs_found) takes it; an env with a plan gets the plan's next open-loop
 * action ([steer in {+1,0,-1}, signed length / step_ratio split into pieces of at most 1]) instead of
 * the policy's.  d_action_out[n][2] is what to pass to hope_step; d_executing[n] = 1 where it came from
 * the plan.  step_ratio = step_len * n_step * VALID_SPEED[1] = 1.25 (train_HOPE_sac.py:164). */
int hope_planner_actions(hope_ctx *ctx, const double *d_policy_action, const hope_out *d_last_out, double *d_action_out,
                         uint8_t *d_executing, double step_ratio, void *stream);
int hope_planner_reset(hope_ctx *ctx, void *stream);

/* Makes `stream` wait until the observation arrays of the last hope_step / hope_reset (lidar, mask, mask_steps, img; pose, target,
 * reward, status, done are complete even earlier) have been written, WITHOUT waiting for the Reeds-Shepp kernels of that step,
 * which hope_step runs next to k_observe.  A rollout loop uses it to start the policy's forward pass for the next action in the
 * shadow of the Reeds-Shepp search (the plan hand-off, which needs rs_*, waits on the stream hope_step was given as usual).
 * HOPE_ERR_INVALID if the last step had no observation stage or was split into several env ranges. */
int hope_wait_observed(hope_ctx *ctx, void *stream);

/* The two non-GEMM steps between the env and the policy network of the rollout loop (row f2), stateless (no context):
 *
 * hope_state_norm — StateNorm.state_norm (model/state_norm.py:25-46) for a batch: with update != 0 the running statistics
 *   d_stats[2][125] (mean then M2 of the 120 lidar + 5 target columns; the caller keeps the sample count and passes the count
 *   BEFORE this batch) take in all n observations (two-pass moments per block, Chan merge), then every observation is normalised as
 *   (x - mean) / (std + 1e-8), std = sqrt(M2 / count), and written as float32; d_mask [n][42] (optional) is cast along.
 *   d_scratch: hope_state_norm_scratch_bytes(n) bytes of device memory.
 * hope_masked_sample — ActionMask.choose_action (model/action_mask.py:199-227) for a batch: d_mean[n][2] float32 policy output
 *   (clamped to [-1,1]), d_log_std[2], d_mask[n][42], d_actions[42][2] = discrete actions in policy scale; draws one action per env
 *   by inverse CDF from a Philox4x32-10 stream keyed by (seed, step, env).  d_index_out / d_u_out (optional): the index drawn
 *   and the uniform variate used. */
int hope_state_norm_scratch_bytes(int n);
int hope_state_norm(const double *d_lidar, const double *d_target, const double *d_mask, int n, double *d_stats, double count_before,
                    int update, void *d_scratch, float *d_out_lidar, float *d_out_target, float *d_out_mask, void *stream);
int hope_masked_sample(int n, const float *d_mean, const double *d_log_std, const double *d_mask, const double *d_actions, uint64_t seed,
                       uint64_t step, double *d_action_out, int32_t *d_index_out, double *d_u_out, void *stream);

/* The actor network between them, as one kernel (row f2 / BASELINE cfg 4): MultiObsEmbedding(ACTOR_CONFIGS) with lidar, target
 * and action-mask inputs (model/network.py:34-196, model/attention.py:16-92, configs.py:134-153) — three 2-layer tanh embeddings,
 * one pre-norm transformer block over the 3 tokens (8 heads x 32, feed-forward 128), Linear(384,128) tanh Linear(128,2) tanh.
 * Weights: every matrix is a bf16 FRAGMENT-PACKED copy of PyTorch's [out][in] weight with the input width zero-padded to a multiple
 * of 16 (120 -> 128, 5 -> 16, 42 -> 48), as hope_policy_pack_matrix produces it (the tensor-core B operand of every 8-row x 16-column
 * tile stored as 256 contiguous bytes, tiles ordered [out / 8][k_pad / 16]: one coalesced load per warp and fragment); biases,
 * LayerNorm parameters and the last layer are float32; all DEVICE pointers.  d_lidar [n][120], d_target [n][5], d_mask [n][42]
 * float32 (hope_state_norm's outputs), d_out [n][2] float32 in [-1, 1].  bf16 operands, float32 accumulation. */
typedef struct hope_policy_weights {
    const void *w1_lidar, *w1_target, *w1_mask; /* packed [128][128], [128][16], [128][48]: embed_*.0.weight, input width padded */
    const void *w2[3];                          /* packed [128][128]: embed_lidar.2 / embed_tgt.2 / embed_am.2 .weight           */
    const void *w_qkv, *w_out;                  /* packed [768][128] attn.fn.to_qkv.weight, [128][256] attn.fn.to_out.0.weight  */
    const void *w_ff1, *w_ff2;                  /* packed [128][128]: ff.fn.net.0 / ff.fn.net.3 .weight                          */
    const void *w_o1;                           /* packed [128][384]: net.output.0.weight                                        */
    const float *b1[3], *b2[3];                 /* [128] each, order lidar, target, mask                                         */
    const float *ln1_g, *ln1_b, *b_out, *ln2_g, *ln2_b, *b_ff1, *b_ff2, *b_o1; /* [128] each                                     */
    const float *w_o2, *b_o2;                   /* [2][128], [2]: net.output.2                                                   */
    const void *w2_img; const float *b2_img;    /* 4-modal network only: packed [128][128] re_embed_img.1.weight, [128] its bias; w_o1 is then packed [128][512] */
} hope_policy_weights;
int hope_policy_forward(int n, const float *d_lidar, const float *d_target, const float *d_mask, const hope_policy_weights *w, float *d_out,
                        void *stream);
/* The 4-modal network (img_shape set, n_modal = 4: what the reference trains with USE_IMG and what its shipped checkpoints hold):
 * d_img_mean [n][128] float32 = embed_img(img)[0], the image encoder's mean head (conv stack by hope_img_conv_forward, then
 * Linear(2048, 256) tanh Linear(256, 128)); the kernel applies re_embed_img (tanh, Linear(128, 128)) as token 3 and runs the
 * 4-token transformer block and the Linear(512, 128) head.  16 envs per CTA. */
int hope_policy_forward_img(int n, const float *d_lidar, const float *d_target, const float *d_mask, const float *d_img_mean,
                            const hope_policy_weights *w, float *d_out, void *stream);
int hope_policy_forward_smem_bytes(void);
/* HOST helper: float32 weight h_w[n_out][n_in] (PyTorch layout) -> the fragment-packed bf16 matrix hope_policy_forward reads,
 * n_out * k_pad bf16 values (n_out multiple of 8, k_pad >= n_in multiple of 16), round to nearest even:
 *   packed[((nt * (k_pad/16) + ks) * 32 + lane) * 4 + 2 * half + e] = W[8 nt + lane / 4][16 ks + 8 half + 2 (lane % 4) + e]. */
int hope_policy_pack_matrix(const float *h_w, int n_out, int n_in, int k_pad, void *h_packed);

/* The convolutional front of the actor's image encoder for the 4-modal network (USE_IMG): ImgEncoder's two residual blocks
 * (model/network.py:198-299 with the shipped switches — no batch norm, tanh, residual): out = maxpool2(tanh(conv3x3(x))) +
 * avgpool2(conv1x1(x)), 3 -> 4 -> 8 channels on d_img / 255, flattened (channel, row, column) = what embed_img.net[0:3]
 * returns.  Weights by VALUE in PyTorch's layouts (float32, HOST memory: they travel in the kernel parameter block);
 * d_img [n][3][64][64] uint8 (hope_outputs.img), d_feat_bf16 [n][2048] bf16.  float32 arithmetic. */
typedef struct hope_img_conv_weights {
    float conv1_w[4 * 3 * 9], conv1_b[4], short1_w[4 * 3], short1_b[4]; /* embed_img.net.0.layer.0 / .shortcut.0 */
    float conv2_w[8 * 4 * 9], conv2_b[8], short2_w[8 * 4], short2_b[8]; /* embed_img.net.1.layer.0 / .shortcut.0 */
} hope_img_conv_weights;
int hope_img_conv_forward(int n, const uint8_t *d_img, const hope_img_conv_weights *w, void *d_feat_bf16, void *stream);

/* State access (device -> host copies; synchronous). */
int hope_get_state(hope_ctx *ctx, double *h_pose, int32_t *h_t, double *h_accum, int32_t *h_scene_id);
int hope_set_state(hope_ctx *ctx, const double *h_pose, const int32_t *h_t, const double *h_accum);
/* counters since creation: [0] env steps with an action, [1] auto-resets, [2] exact-orientation
 * fallbacks taken, [3] RS word-capacity overflows, [4] RS zero-length words (reference asserts),
 * [5] kernels launched by this context, [6] scenes generated on the device */
int hope_get_counters(hope_ctx *ctx, uint64_t h_counters[8]);
/* Per-kernel device timing: when enabled (on = 1) every kernel launch of a step is bracketed by CUDA
 * events on the launch stream; on = 2 additionally keeps all kernels of hope_step on ONE stream (k_observe does not overlap
 * the Reeds-Shepp chain), so each duration is that kernel running alone inside the live loop.  hope_profile_read synchronises, returns the accumulated
 * milliseconds and launch counts per kernel [0] advance [1] observe [2] rs_enumerate [3] rs_walk
 * [4] rs_check [5] rs_select [6] render ([7] reserved) since the last read, and clears them. */
int hope_profile_enable(hope_ctx *ctx, int on);
int hope_profile_read(hope_ctx *ctx, double h_ms[8], uint64_t h_launches[8]);
/* Micro-benchmark for the roofline's compute axis: sustained float64 FMA rate of the device (8 independent DFMA
 * chains per thread, SMs x 8 blocks x 256 threads), in TFLOP/s counting an FMA as 2.  Synchronous, ~20 ms. */
int hope_fp64_peak_tflops(int device, double *tflops);
int hope_n_envs(const hope_ctx *ctx);
int hope_max_obs(void);
int hope_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HOPE_B200_H */
