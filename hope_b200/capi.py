"""ctypes binding of include/hope_b200.h (the C ABI of libhope_b200.so)."""
import ctypes as C
import os

from . import build as _build

MAX_OBS, MAX_VERTS, N_LIDAR, N_ACTION, N_MASK_ITER, N_UPSAMPLE, RS_MAX_SEG = 16, 4, 120, 42, 10, 1200, 5
STAGE_ADVANCE, STAGE_OBSERVE, STAGE_RS, STAGE_ALL, STAGE_IMAGE, STAGE_RAW_ACTION = 1, 2, 4, 7, 8, 16
IMG_C, IMG_HW, N_COLOR = 3, 64, 25
CONTINUE, ARRIVED, COLLIDED, OUTBOUND, OUTTIME = 1, 2, 3, 4, 5
RS_S, RS_L, RS_R, RS_NONE = 0, 1, 2, 255


class HopeError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("wheel_base", C.c_double), ("box_x", C.c_double * 4), ("box_y", C.c_double * 4),
        ("valid_speed", C.c_double * 2), ("valid_steer", C.c_double * 2), ("num_step", C.c_int),
        ("step_length", C.c_double), ("mini_iter", C.c_int), ("lidar_range", C.c_double),
        ("tolerant_time", C.c_int), ("rs_max_dist", C.c_double), ("rs_step", C.c_double),
        ("reward_weight", C.c_double * 5), ("reward_ratio", C.c_double), ("env_collide", C.c_int),
        ("auto_reset", C.c_int), ("regen_on_reset", C.c_int), ("regen_level", C.c_int), ("regen_seed", C.c_uint64)]


# name -> (ctype, trailing shape); order must match struct hope_out
OUT_FIELDS = [
    ("pose", C.c_double, (3,)), ("lidar", C.c_double, (N_LIDAR,)), ("mask", C.c_double, (N_ACTION,)),
    ("mask_steps", C.c_uint8, (N_ACTION,)), ("target", C.c_double, (5,)), ("reward", C.c_double, ()),
    ("reward_info", C.c_double, (5,)), ("status", C.c_int32, ()), ("done", C.c_uint8, ()),
    ("substeps", C.c_uint8, ()), ("retreated", C.c_uint8, ()), ("was_reset", C.c_uint8, ()),
    ("rs_found", C.c_uint8, ()), ("rs_nseg", C.c_uint8, ()), ("rs_types", C.c_uint8, (RS_MAX_SEG,)),
    ("rs_lengths", C.c_double, (RS_MAX_SEG,)), ("rs_L", C.c_double, ()), ("rs_ncand", C.c_uint8, ()),
    ("rs_ntried", C.c_uint8, ()), ("img", C.c_uint8, (IMG_C, IMG_HW, IMG_HW))]


class Out(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in OUT_FIELDS]


_LIBS = {}


class PolicyWeights(C.Structure):
    """hope_policy_weights (include/hope_b200.h): device pointers of the actor network's parameters"""
    _fields_ = ([("w1_lidar", C.c_void_p), ("w1_target", C.c_void_p), ("w1_mask", C.c_void_p), ("w2", C.c_void_p * 3),
                 ("w_qkv", C.c_void_p), ("w_out", C.c_void_p), ("w_ff1", C.c_void_p), ("w_ff2", C.c_void_p), ("w_o1", C.c_void_p),
                 ("b1", C.c_void_p * 3), ("b2", C.c_void_p * 3)] +
                [(k, C.c_void_p) for k in ("ln1_g", "ln1_b", "b_out", "ln2_g", "ln2_b", "b_ff1", "b_ff2", "b_o1", "w_o2", "b_o2", "w2_img", "b2_img")])


class ImgConvWeights(C.Structure):
    """hope_img_conv_weights (include/hope_b200.h): the two residual conv blocks of the image encoder, by value, PyTorch layouts"""
    _fields_ = [("conv1_w", C.c_float * 108), ("conv1_b", C.c_float * 4), ("short1_w", C.c_float * 12), ("short1_b", C.c_float * 4),
                ("conv2_w", C.c_float * 288), ("conv2_b", C.c_float * 8), ("short2_w", C.c_float * 32), ("short2_b", C.c_float * 8)]


def load_library(max_obs=16):
    """Load (building if necessary) libhope_b200.so, or its 128-obstacle build for max_obs=128.  There is no
    fallback path: a missing toolchain or library is an error."""
    if max_obs in _LIBS:
        return _LIBS[max_obs]
    if max_obs not in _build.VARIANTS:
        raise HopeError(f"no build with {max_obs} obstacle rings per scene (have {sorted(_build.VARIANTS)})")
    path = _build.VARIANTS[max_obs]
    if _build.needs_build(max_obs):
        path = _build.build(variants=(max_obs,))
    if not os.path.exists(path):
        raise HopeError(f"{path} is missing; run `python -m hope_b200.build --all`")
    lib = C.CDLL(path)
    vp, i32, u64, dp, ip = C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p
    sig = {
        "hope_version": (C.c_int, []),
        "hope_default_params": (C.c_int, [C.POINTER(Params)]),
        "hope_create": (C.c_int, [C.POINTER(vp), i32, i32, i32, C.POINTER(Params)]),
        "hope_destroy": (C.c_int, [vp]),
        "hope_strerror": (C.c_char_p, [i32]),
        "hope_last_cuda_error": (C.c_char_p, [vp]),
        "hope_upload_tables": (C.c_int, [vp] + [dp] * 7),
        "hope_set_palette": (C.c_int, [vp, vp]),
        "hope_set_render_traj": (C.c_int, [vp, C.c_int]),
        "hope_expand_mask": (C.c_int, [vp, vp, i32]),
        "hope_expand_mask_portable": (C.c_int, [vp, vp, i32]),
        "hope_expand_lidar": (C.c_int, [vp, vp, vp, vp, vp, i32, i32]),
        "hope_host_wire_info": (C.c_int, [vp, C.POINTER(u64 * 8)]),
        "hope_set_scene_pool": (C.c_int, [vp, i32, i32, dp, dp, dp, dp, ip]),
        "hope_generate_scenes": (C.c_int, [i32, i32, u64, i32, dp, dp, dp, dp, ip, ip]),
        "hope_generate_scene_pool_device": (C.c_int, [vp, i32, i32, i32, u64, vp]),
        "hope_get_scene_pool": (C.c_int, [vp, i32, i32, dp, dp, dp, dp, ip]),
        "hope_reset": (C.c_int, [vp, ip, C.POINTER(Out), vp]),
        "hope_step": (C.c_int, [vp, dp, C.POINTER(Out), C.c_uint, vp]),
        "hope_step_kinematics_collision": (C.c_int, [vp, dp, dp, vp, vp, vp]),
        "hope_step_host": (C.c_int, [vp, dp, C.POINTER(Out), C.c_uint]),
        "hope_reset_host": (C.c_int, [vp, ip, C.POINTER(Out)]),
        "hope_get_state": (C.c_int, [vp, dp, ip, dp, ip]),
        "hope_set_state": (C.c_int, [vp, dp, ip, dp]),
        "hope_get_counters": (C.c_int, [vp, C.POINTER(u64 * 8)]),
        "hope_n_envs": (C.c_int, [vp]),
        "hope_max_obs": (C.c_int, []),
        "hope_planner_actions": (C.c_int, [vp, dp, C.POINTER(Out), dp, vp, C.c_double, vp]),
        "hope_planner_reset": (C.c_int, [vp, vp]),
        "hope_wait_observed": (C.c_int, [vp, vp]),
        "hope_fp64_peak_tflops": (C.c_int, [i32, C.POINTER(C.c_double)]),
        "hope_state_norm_scratch_bytes": (C.c_int, [i32]),
        "hope_state_norm": (C.c_int, [dp, dp, dp, i32, dp, C.c_double, i32, vp, vp, vp, vp, vp]),
        "hope_masked_sample": (C.c_int, [i32, vp, dp, dp, dp, u64, u64, dp, vp, vp, vp]),
        "hope_policy_forward": (C.c_int, [i32, vp, vp, vp, C.POINTER(PolicyWeights), vp, vp]),
        "hope_policy_forward_img": (C.c_int, [i32, vp, vp, vp, vp, C.POINTER(PolicyWeights), vp, vp]),
        "hope_policy_forward_smem_bytes": (C.c_int, []),
        "hope_policy_pack_matrix": (C.c_int, [vp, i32, i32, i32, vp]),
        "hope_img_conv_forward": (C.c_int, [i32, vp, C.POINTER(ImgConvWeights), vp, vp]),
        "hope_profile_enable": (C.c_int, [vp, i32]),
        "hope_profile_read": (C.c_int, [vp, C.POINTER(C.c_double * 8), C.POINTER(u64 * 8)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here means the library does not match the header
        fn.restype, fn.argtypes = res, args
    if lib.hope_max_obs() != max_obs:
        raise HopeError(f"{path} was built for {lib.hope_max_obs()} obstacle rings, expected {max_obs}")
    _LIBS[max_obs] = lib
    return lib


def check(rc, ctx=None):
    if rc == 0:
        return
    lib = load_library()
    msg = lib.hope_strerror(rc).decode()
    if ctx is not None and rc in (-2, -5) and lib.hope_last_cuda_error(ctx):
        msg += ": " + lib.hope_last_cuda_error(ctx).decode()
    raise HopeError(f"hope_b200 error {rc}: {msg}")
