"""BatchedParkingEnv — N independent parking scenes stepped on one B200 through the C ABI.

Batched counterpart of the reference's single-env surface
    CarParkingWrapper.reset / .step     src/env/env_wrapper.py:58-85
    CarParking.reset / .step            src/env/car_parking_base.py:127-138, 235-299
Observation keys, dtypes (float64) and layouts are the reference's, with a leading env axis:
    obs['lidar'] (N,120)  obs['target'] (N,5)  obs['action_mask'] (N,42)
    obs['img'] None, or with use_img_observation=True the (N,3,64,64) uint8 image whose /255.0 is the reference's
    float64 image (car_parking_base.py:301-350, observation_processor.py:13-23; 12 KB instead of 96 KB per env)
`action_mask[:, j]`: j<21 forward, j>=21 backward, steer 0.75-0.075*(j%21)  (configs.py:108-115).

PyTorch is used for device memory and streams only; all compute is in libhope_b200.so.  There
is no CPU fallback: constructing the env without CUDA raises.
"""
import ctypes as C

import numpy as np

from . import capi, refconfig, tables

LEVELS = {"Normal": 0, "Complex": 1, "Extrem": 2}
_NP = {C.c_double: np.float64, C.c_uint8: np.uint8, C.c_int32: np.int32}


def generate_scenes(n, level="Normal", seed=42, nthreads=0, max_obs=16):
    """Host-side procedural scenes (parking_map_normal.py:40-494 semantics, own RNG streams).
    level: 'Normal' | 'Complex' | 'Extrem' | 'mix' (i % 3 cycles the three, BASELINE cfg 3)."""
    lib = capi.load_library(max_obs)
    if level == "mix":
        parts = [generate_scenes((n - k + 2) // 3, lv, seed + 1000003 * k, nthreads, max_obs) for k, lv in enumerate(LEVELS)]
        out = {key: np.zeros((n,) + parts[0][key].shape[1:], dtype=parts[0][key].dtype) for key in parts[0]}
        for k in range(3):
            for key in out:
                out[key][k::3] = parts[k][key]
        return out
    sc = dict(start=np.zeros((n, 3)), dest=np.zeros((n, 3)), bounds=np.zeros((n, 4)),
              obs=np.zeros((n, max_obs, capi.MAX_VERTS, 2)), nverts=np.zeros((n, max_obs), dtype=np.int32),
              case_id=np.zeros(n, dtype=np.int32))
    capi.check(lib.hope_generate_scenes(n, LEVELS[level], seed, nthreads, sc["start"].ctypes.data, sc["dest"].ctypes.data,
                                        sc["bounds"].ctypes.data, sc["obs"].ctypes.data, sc["nverts"].ctypes.data,
                                        sc["case_id"].ctypes.data))
    return sc


class BatchedParkingEnv(object):
    def __init__(self, n_envs, scenes=None, pool_size=None, level="Normal", seed=42, device=0, auto_reset=True,
                 params=None, device_scenes=False, max_obs=None, use_img_observation=False, config=None):
        """scenes: dict of host arrays (start/dest/bounds/obs/nverts) to upload as the pool; None -> generate
        `pool_size` (default 2 n) scenes of `level` on the host, or with device_scenes=True on the GPU
        (then finished envs also get a fresh device-generated scene instead of cycling the pool).
        config: a configs snapshot (refconfig.from_module / defaults); None -> refconfig.load(), i.e. the caller's own
        `configs` module when one is importable, so that edits to the reference's configs.py keep taking effect.
        params: hope_params fields that override the snapshot (e.g. {'tolerant_time': 50})."""
        import torch
        if not torch.cuda.is_available():
            raise capi.HopeError("BatchedParkingEnv needs a CUDA device; there is no CPU path")
        self.torch = torch
        if max_obs is None:  # scenes decide: anything wider than the default block needs the 128-ring build
            max_obs = int(scenes["nverts"].shape[1]) if scenes is not None else 16
        self.max_obs = max_obs
        self.lib = capi.load_library(max_obs)
        self.n = int(n_envs)
        self.device = torch.device("cuda", device)
        if scenes is None and not device_scenes:
            pool_size = pool_size or 2 * self.n
            scenes = generate_scenes(pool_size, level, seed, max_obs=max_obs)
        self.scenes = scenes
        self.pool_size = int(scenes["start"].shape[0]) if scenes is not None else int(pool_size or self.n)
        self.config = refconfig.validate(config or refconfig.load())
        self.params = capi.Params()
        capi.check(self.lib.hope_default_params(C.byref(self.params)))
        for k, v in refconfig.step_params(self.config).items():
            if isinstance(v, list):
                arr = getattr(self.params, k)
                for q, x in enumerate(v):
                    arr[q] = x
            else:
                setattr(self.params, k, v)
        self.params.auto_reset = 1 if auto_reset else 0
        if device_scenes:
            self.params.regen_on_reset = 1
            self.params.regen_level = -1 if level == "mix" else LEVELS[level]
            self.params.regen_seed = seed
        for k, v in (params or {}).items():
            setattr(self.params, k, v)
        self.ctx = C.c_void_p()
        capi.check(self.lib.hope_create(C.byref(self.ctx), device, self.n, self.pool_size, C.byref(self.params)), self.ctx)
        tb = tables.host_tables(self.config)
        capi.check(self.lib.hope_upload_tables(self.ctx, *[tb[k].ctypes.data for k in
                                                            ("ray_a", "ray_b", "lidar_base", "mask_base", "dist_star", "w_lo", "w_hi")]), self.ctx)
        if scenes is not None:
            self.set_scene_pool(scenes)
        else:
            self.generate_pool_on_device(0, self.pool_size, level, seed)
        # device outputs (torch owns the memory; the library only sees raw pointers)
        self.out = {}
        self._out_struct = capi.Out()
        self.use_img = bool(use_img_observation)
        if self.use_img:
            if not refconfig.image_supported(self.config):
                raise capi.HopeError("image observation: k_render is compiled for WIN 500x500, OBS 256x256, K 12, TRAJ_RENDER_LEN <= 20 "
                                     "(car_parking_base.py:301-350); configs.py asks for another raster geometry")
            self.set_palette(refconfig.palette(self.config))
            capi.check(self.lib.hope_set_render_traj(self.ctx, refconfig.traj_render_len(self.config)), self.ctx)
        self.default_stages = capi.STAGE_ALL | (capi.STAGE_IMAGE if self.use_img else 0)
        for name, ct, shape in capi.OUT_FIELDS:
            if name == "img" and not self.use_img:
                continue  # NULL pointer: the library skips the image stage
            tdt = {C.c_double: torch.float64, C.c_uint8: torch.uint8, C.c_int32: torch.int32}[ct]
            t = torch.zeros((self.n,) + shape, dtype=tdt, device=self.device)
            self.out[name] = t
            setattr(self._out_struct, name, t.data_ptr())
        self._host = None

    # ------------------------------------------------------------------------------------------
    def set_scene_pool(self, scenes, first=0):
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        s, d, b, o = f8(scenes["start"]), f8(scenes["dest"]), f8(scenes["bounds"]), f8(scenes["obs"])
        nv = np.ascontiguousarray(scenes["nverts"], dtype=np.int32)
        if nv.shape[1] != self.max_obs or o.shape[1] != self.max_obs:
            raise capi.HopeError(f"scene arrays hold {nv.shape[1]} rings per scene, this env was built for {self.max_obs}")
        capi.check(self.lib.hope_set_scene_pool(self.ctx, first, s.shape[0], s.ctypes.data, d.ctypes.data, b.ctypes.data,
                                                o.ctypes.data, nv.ctypes.data), self.ctx)

    def generate_pool_on_device(self, first, n, level="mix", seed=42):
        """hope_generate_scene_pool_device: fill pool slots [first, first+n) with scenes generated by GPU threads."""
        lv = -1 if level == "mix" else LEVELS[level]
        capi.check(self.lib.hope_generate_scene_pool_device(self.ctx, first, n, lv, seed, self._stream()), self.ctx)

    def get_scene_pool(self, first=0, n=None):
        """Read pool scenes back to the host (e.g. to hand device-generated scenes to a checker)."""
        n = self.pool_size - first if n is None else n
        sc = dict(start=np.zeros((n, 3)), dest=np.zeros((n, 3)), bounds=np.zeros((n, 4)),
                  obs=np.zeros((n, self.max_obs, capi.MAX_VERTS, 2)), nverts=np.zeros((n, self.max_obs), dtype=np.int32))
        capi.check(self.lib.hope_get_scene_pool(self.ctx, first, n, sc["start"].ctypes.data, sc["dest"].ctypes.data, sc["bounds"].ctypes.data,
                                                sc["obs"].ctypes.data, sc["nverts"].ctypes.data), self.ctx)
        return sc

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _obs(self):
        return {"img": self.out.get("img"), "lidar": self.out["lidar"], "target": self.out["target"], "action_mask": self.out["mask"]}

    def _info(self):
        o = self.out
        return {"status": o["status"], "reward_info": o["reward_info"], "was_reset": o["was_reset"],
                "path_to_dest": {"found": o["rs_found"], "nseg": o["rs_nseg"], "types": o["rs_types"],
                                 "lengths": o["rs_lengths"], "L": o["rs_L"]}}

    # ---- device-tensor API (what a GPU-resident rollout loop calls; asynchronous) ------------------
    def reset(self, scene_ids=None):
        ids = None
        if scene_ids is not None:
            ids = np.ascontiguousarray(scene_ids, dtype=np.int32)
        capi.check(self.lib.hope_reset(self.ctx, ids.ctypes.data if ids is not None else None, C.byref(self._out_struct),
                                       self._stream()), self.ctx)
        return self._obs()

    def step(self, actions, stages=None, raw_action=False):
        """actions: (N,2) float64 CUDA tensor in [-1,1] (steer, speed), as the policy emits them (raw_action=True: physical
        [steer rad, speed m/s] as CarParking.step takes them); None = a step without motion (CarParking.step(None))."""
        t = self.torch
        stages = self.default_stages if stages is None else stages
        if raw_action:
            stages |= capi.STAGE_RAW_ACTION
        if actions is not None and not (isinstance(actions, t.Tensor) and actions.is_cuda and actions.dtype == t.float64 and actions.is_contiguous()
                                        and tuple(actions.shape) == (self.n, 2)):
            raise capi.HopeError("step() expects a contiguous float64 CUDA tensor of shape (n_envs, 2)")
        capi.check(self.lib.hope_step(self.ctx, actions.data_ptr() if actions is not None else None, C.byref(self._out_struct), stages,
                                      self._stream()), self.ctx)
        return self._obs(), self.out["reward"], self.out["done"], self._info()

    def wait_observed(self, stream):
        """Make the torch stream `stream` wait for the observation of the last step() / reset() without waiting for that step's
        Reeds-Shepp kernels (hope_wait_observed).  Returns False when the library cannot offer that (the caller then orders
        itself behind the stream step() ran on, as usual)."""
        return self.lib.hope_wait_observed(self.ctx, stream.cuda_stream) == 0

    def step_kinematics_collision(self, actions):
        """BASELINE cfg 2: pose integration + collision (+ arrival) only."""
        o = self.out
        capi.check(self.lib.hope_step_kinematics_collision(self.ctx, actions.data_ptr(), o["pose"].data_ptr(),
                                                           o["retreated"].data_ptr(), o["substeps"].data_ptr(), self._stream()), self.ctx)
        return o["pose"], o["retreated"], o["substeps"]

    # ---- host-buffer API (what the reference's host-side training loop would call; synchronous) ----
    HOST_DEFAULT = ("lidar", "mask", "target", "reward", "done", "status", "reward_info", "rs_found", "rs_nseg",
                    "rs_types", "rs_lengths")

    def _host_buffers(self, names):
        t = self.torch
        if self._host is None:
            self._host = {"action": t.zeros((self.n, 2), dtype=t.float64).pin_memory()}
        st = capi.Out()
        for name, ct, shape in capi.OUT_FIELDS:
            if name not in names:
                continue
            if name not in self._host:
                tdt = {C.c_double: t.float64, C.c_uint8: t.uint8, C.c_int32: t.int32}[ct]
                self._host[name] = t.zeros((self.n,) + shape, dtype=tdt).pin_memory()
            setattr(st, name, self._host[name].data_ptr())
        return st

    def step_host(self, actions, stages=None, outputs=HOST_DEFAULT, raw_action=False):
        """actions: (N,2) float64 numpy array on the host: the policy's output in [-1,1]^2 (CarParkingWrapper.step), or with
        raw_action=True the physical [steer, speed] CarParking.step takes; None = a step without motion (CarParking.step(None)).
        Copies it in, steps, copies `outputs` back into pinned host buffers, synchronises.  Returns dict of numpy views."""
        st = self._host_buffers(outputs)
        stages = (capi.STAGE_ALL | (capi.STAGE_IMAGE if "img" in outputs else 0)) if stages is None else stages
        if raw_action:
            stages |= capi.STAGE_RAW_ACTION
        ptr = None
        if actions is not None:
            self._host["action"].numpy()[...] = actions
            ptr = self._host["action"].data_ptr()
        capi.check(self.lib.hope_step_host(self.ctx, ptr, C.byref(st), stages), self.ctx)
        return {k: self._host[k].numpy() for k in outputs}

    def reset_host(self, scene_ids=None, outputs=HOST_DEFAULT):
        st = self._host_buffers(outputs)
        ids = None
        if scene_ids is not None:
            ids = np.ascontiguousarray(scene_ids, dtype=np.int32)
        capi.check(self.lib.hope_reset_host(self.ctx, ids.ctypes.data if ids is not None else None, C.byref(st)), self.ctx)
        return {k: self._host[k].numpy() for k in outputs}

    def host_io_bytes(self):
        """(h2d, d2h) bytes the LAST step_host call moved over PCIe, counted by the library from the copies it issued.  The
        float64 action mask travels as its 42 uint8 step counts and the lidar as flag bits + the beams that hit something;
        host threads inside hope_step_host rebuild both (HOPE_B200_HOST_MASK_EXPAND=0 / HOPE_B200_HOST_LIDAR_PACK=0 turn it off)."""
        w = self.host_wire_info()
        return w["h2d_bytes"], w["d2h_bytes"]

    def host_wire_info(self):
        info = (C.c_uint64 * 8)()
        capi.check(self.lib.hope_host_wire_info(self.ctx, C.byref(info)), self.ctx)
        return {"h2d_bytes": int(info[0]), "d2h_bytes": int(info[1]), "mask_narrow": bool(info[2]), "lidar_packed": bool(info[3]),
                "host_threads": int(info[4]), "avx512": bool(info[5]), "env_ranges": int(info[6]), "lidar_packed_envs": int(info[7])}

    # ---- state / diagnostics -------------------------------------------------------------------
    def get_state(self):
        pose = np.zeros((self.n, 3)); t = np.zeros(self.n, dtype=np.int32); acc = np.zeros(self.n); sid = np.zeros(self.n, dtype=np.int32)
        capi.check(self.lib.hope_get_state(self.ctx, pose.ctypes.data, t.ctypes.data, acc.ctypes.data, sid.ctypes.data), self.ctx)
        return dict(pose=pose, t=t, accum=acc, scene_id=sid)

    def set_state(self, pose=None, t=None, accum=None):
        p = np.ascontiguousarray(pose, dtype=np.float64) if pose is not None else None
        tt = np.ascontiguousarray(t, dtype=np.int32) if t is not None else None
        a = np.ascontiguousarray(accum, dtype=np.float64) if accum is not None else None
        capi.check(self.lib.hope_set_state(self.ctx, p.ctypes.data if p is not None else None, tt.ctypes.data if tt is not None else None,
                                           a.ctypes.data if a is not None else None), self.ctx)

    def counters(self):
        buf = (C.c_uint64 * 8)()
        capi.check(self.lib.hope_get_counters(self.ctx, C.byref(buf)), self.ctx)
        names = ("env_steps", "auto_resets", "exact_orient_fallbacks", "rs_capacity_overflows", "rs_zero_length_words", "kernel_launches",
                 "device_scenes_generated")
        return {k: int(buf[i]) for i, k in enumerate(names)}

    # ---- batched RsPlanner hand-off (parking_agent.py:2-47) ----------------------------------------
    def planner_actions(self, policy_actions, step_ratio=1.25):
        """Replace the policy's action by the RS plan's next open-loop action where a plan is being
        executed.  Uses the outputs of the previous step; returns (actions, executing) device tensors."""
        t = self.torch
        if not hasattr(self, "_plan_action"):
            self._plan_action = t.zeros((self.n, 2), dtype=t.float64, device=self.device)
            self._plan_exec = t.zeros(self.n, dtype=t.uint8, device=self.device)
        capi.check(self.lib.hope_planner_actions(self.ctx, policy_actions.data_ptr(), C.byref(self._out_struct), self._plan_action.data_ptr(),
                                                 self._plan_exec.data_ptr(), float(step_ratio), self._stream()), self.ctx)
        return self._plan_action, self._plan_exec

    def planner_reset(self):
        capi.check(self.lib.hope_planner_reset(self.ctx, self._stream()), self.ctx)

    KERNELS = ("k_advance", "k_observe", "k_rs_enumerate", "k_rs_walk", "k_rs_check", "k_rs_select", "k_render")

    def set_palette(self, rgb):
        """hope_set_palette: (25,3) uint8 colours in painter's order (configs.py:26-30, 80-88)."""
        a = np.ascontiguousarray(rgb, dtype=np.uint8)
        if a.shape != (capi.N_COLOR, 3):
            raise capi.HopeError("palette must be (25, 3) uint8")
        capi.check(self.lib.hope_set_palette(self.ctx, a.ctypes.data), self.ctx)

    def profile(self, on=True, serial=False):
        """per-kernel CUDA-event timing; serial=True keeps the whole step on one stream so every kernel is timed running alone"""
        capi.check(self.lib.hope_profile_enable(self.ctx, (2 if serial else 1) if on else 0), self.ctx)

    def profile_read(self):
        """{kernel: (total_ms, launches)} measured with CUDA events on the launch stream."""
        ms = (C.c_double * 8)(); cnt = (C.c_uint64 * 8)()
        capi.check(self.lib.hope_profile_read(self.ctx, C.byref(ms), C.byref(cnt)), self.ctx)
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    def close(self):
        if self.ctx:
            self.lib.hope_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
