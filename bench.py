#!/usr/bin/env python
"""bench.py — env-steps/s of the batched ParkingEnv step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (C oracle on the host cores; never loads the product library)

Workload = BASELINE.json configs[2]: 65 536 parallel scenes per GPU (levels Normal/Complex/Extrem
cycled), FULL step: kinematics + ring collision + arrival + status/reward, 120-beam LiDAR raycast,
42-action mask sweep, Reeds-Shepp search; float64 random actions U(-1,1)^2; finished envs take
their next scene from a pre-generated pool of 2N scenes on the following step (those reset
steps are NOT counted as env-steps).  One "step" = one pass over all scenes of the rank.

Timed regions of the main line
  value      device-resident: actions already in HBM, hope_step on the current stream, CUDA events, max over ranks.
  e2e        hope_step_host: pinned host actions -> H2D, step, D2H of the observation/reward/done/RS buffers, synchronised;
             wall clock around the synchronous calls, max over ranks.
  roofline   per-kernel CUDA-event durations of a second, SERIALISED pass over the same steps (every kernel alone on one
             stream): the dominant kernel and its HBM fraction come from these, not from the overlapped live step.
  secondary  BASELINE cfg 4 (PPO acting loop with the transformer policy) and cfg 5 (SAC rollout + update + NCCL gradient
             all-reduce) on the same envs, a few steps each, with their own clock records; cfg4_img = cfg 4 as the reference ships
             it (USE_IMG: image observation rendered every step, 4-modal actor).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
WORKLOAD = "cfg3: 65536 scenes/GPU, full step (advance+collision, lidar raycast, mask sweep, Reeds-Shepp), levels mixed"
# Algorithmic bytes per env-step (SURVEY.md §8d / BASELINE.md §4; fp64 state and observations as the
# reference emits them, 6.1 quads = 390 B of vertices, every datum moved once):
ALGO_BYTES = {
    "k_advance": 516,              # pose 24 r/w + action 16 + vertices 390 + n_obs 4 + dest/bounds 56 + flags 2
    "k_observe": 1378 + 62 + 336 + 40,  # raycast (pose, vertices, lidar 960 out) + table share + mask f64 + target; lidar not re-read (fused)
    "k_rs_enumerate": 24 + 24 + 16,     # pose + dest + gate/state in; word list stays on chip-side scratch
    "k_rs_walk": 0.75 * 2.9 * (56 + 560),   # per gated env-step: ~2.9 tried words, 56 B word in + 560 B sampling plan out
    "k_rs_check": 24 + 56 + 394 + 0.75 * 2.9 * 560,   # pose, dest/bounds, vertices + the sampling plans read back
    "k_rs_select": 47,                   # RS result out
    "k_render": 12288 + 24 + 394 + 16 + 64 + 484,  # image out (3x64x64 u8) + pose, vertices, cs, bounds/start/dest, 20-pose trajectory tail
}
FULL_STEP_BYTES = 1986  # SURVEY.md §8d, whole fused step
CPU_SAMPLE_STEPS = 8     # cpu_baseline leg of the main arm: 65 536 scenes x 8 steps, ~10 s on 16 host threads
ONE_CORE_ENVS, ONE_CORE_STEPS = 1024, 4
PY_REFERENCE_NOTE = ("the reference's own single-process Python env cannot run on the GPU box (shapely/gym/pygame/heapdict are not installed and the "
                     "reference tree is absent); measured in the build container with stand-ins for those packages: 43-54 env-steps/s on one core "
                     "(profiles/r01_reference_python_timing.json), 65-87 in the survey probe (BASELINE.md §2)")


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def scene_seed(rank):
    """scene id -> GPU: rank r generates its own pool of the global synthetic scene set."""
    return 42 + 7919 * rank


def reduce_scalar(x, op, world, device):
    """max / sum of a python float over ranks (the only collectives of the env path)."""
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def host_threads_for_rank(world):
    """host threads one rank may use (mask expansion inside hope_step_host): the box's CPUs shared evenly"""
    cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(12, cpus // max(1, world) - 1))


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first_line(self, keep_busy, limit=3.0):
        """let nvidia-smi deliver its first line while `keep_busy()` keeps the GPU under load (untimed, uncounted)"""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < limit:
            keep_busy()
        self.lines.clear()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, flag in zip(names, p[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def timed_with_clocks(rank, local_rank, world, dev, run, keep_busy, busy_frac, min_seconds=0.4):
    """CUDA-event time of `run()` (max over ranks) with nvidia-smi sampling SM clocks / throttle reasons during the region.
    A region shorter than `min_seconds` is followed by untimed repeats of `keep_busy` (each worth `busy_frac` of the region) so
    the sampler sees the clocks under the same load; every rank runs the same number of repeats (keep_busy may contain
    collectives), they are outside the events, and the callers subtract their env-steps."""
    import math
    import torch
    import torch.distributed as dist
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world == 1:
        sampler.wait_first_line(keep_busy)
    else:
        for _ in range(2):
            keep_busy()
        if rank == 0:
            sampler.lines.clear()
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    min_ms = 1e3 * min_seconds * (1.0 if world == 1 else 1.5)
    reps = 0 if ms >= min_ms else int(math.ceil((min_ms - ms) / max(ms * busy_frac, 1e-3)))
    for _ in range(reps):
        keep_busy()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    return ms, clocks


# ---- CPU arm: the C oracle, never the product --------------------------------------------------------------------------------
def fixture_scenes(n):
    """n scenes for the CPU arm WITHOUT the product library: the 288 scenes recorded from the reference's own generator
    (tests/golden/scenes_{Normal,Complex,Extrem}.npz, 96 per level), levels cycled like the GPU arm, tiled to n."""
    parts = [dict(np.load(os.path.join(ROOT, "tests", "golden", f"scenes_{lv}.npz"))) for lv in ("Normal", "Complex", "Extrem")]
    k = len(parts[0]["start"])
    idx = np.arange(n)
    out = {}
    for key in ("start", "dest", "bounds", "obs", "nverts"):
        stack = np.stack([p[key] for p in parts])          # [3][96]...
        out[key] = np.ascontiguousarray(stack[idx % 3, (idx // 3) % k])
    return out


def cpu_oracle_rate(n_envs, steps, nthreads, warmup=1, seed=42):
    """env-steps/s of the C oracle (oracle/c/parking_oracle.c, OpenMP over scenes) over `steps` passes of n_envs scenes of
    the cfg-3 workload; auto-reset on the following step and reset steps not counted, like the GPU arm."""
    from oracle import parking_oracle as po
    sc = fixture_scenes(n_envs)
    env = po.OracleEnv(sc["start"], sc["dest"], sc["bounds"], sc["obs"], sc["nverts"], nthreads=nthreads)
    env.reset_step()
    rng = np.random.default_rng(seed)
    state = {"done": np.zeros(n_envs, dtype=bool)}

    def one():
        done = state["done"]
        act = rng.uniform(-1, 1, size=(n_envs, 2))
        if done.any():
            env.reset_state(np.where(done)[0])
        out = env.step(act, has_action=(~done).astype(np.uint8))
        n = int((~done).sum())
        state["done"] = out["status"] != 1
        return n

    for _ in range(max(1, warmup)):  # pages in the 4 MB table, spreads the episodes over their lengths
        one()
    counted, t0 = 0, time.perf_counter()
    for _ in range(steps):
        counted += one()
    dt = time.perf_counter() - t0
    return counted / dt, counted, dt


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.envs
    rate, counted, dt = cpu_oracle_rate(n, max(1, args.steps), threads, warmup=max(1, args.warmup))
    one_rate, one_counted, one_dt = cpu_oracle_rate(ONE_CORE_ENVS, ONE_CORE_STEPS, 1)
    assert "hope_b200" not in sys.modules, "the reference arm must not depend on the product library"
    sample = f"C oracle (oracle/c/parking_oracle.c, OpenMP x{threads}) on {n} scenes x {args.steps} steps = {counted} env-steps in {dt:.1f} s"
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": n,
                   "scenes": "288 scenes recorded from the reference's own generator (tests/golden/scenes_*.npz), levels cycled, tiled to the env count; "
                             "every env has its own action stream",
                   "sample": f"{n} scenes x {args.steps} steps per run (one step = one pass over all scenes, as on the GPU arm)"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample,
                         "one_core": {"value": one_rate, "unit": "env-steps/s", "cores": 1,
                                      "sample": f"same C oracle, 1 thread, {ONE_CORE_ENVS} scenes x {ONE_CORE_STEPS} steps = {one_counted} env-steps in {one_dt:.1f} s"},
                         "python_reference": PY_REFERENCE_NOTE},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- secondary configurations ---------------------------------------------------------------------------------------------------
def measure_rollout(env, rank, local_rank, world, dev, steps, warmup, use_img=False):
    """BASELINE cfg 4: PPO acting loop: state norm -> transformer policy (bf16 autocast, tensor-core GEMMs through stock PyTorch)
    -> masked discrete sampling -> RS plan hand-off -> env step; device-resident."""
    import torch
    from hope_b200 import rollout
    actor, what = rollout.reference_actor(use_img=use_img, device=dev)
    eng = rollout.RolloutEngine(env, actor, seed=rank)
    eng.collect(warmup)
    torch.cuda.synchronize()
    c0 = env.counters()
    ran = {"n": 0}

    def run():
        eng.collect(steps); ran["n"] += steps

    def busy():
        eng.collect(2); ran["n"] += 2

    ms, clocks = timed_with_clocks(rank, local_rank, world, dev, run, busy, 2.0 / steps)
    c1 = env.counters()
    counted = reduce_scalar((c1["env_steps"] - c0["env_steps"]) * steps / max(1, ran["n"]), "sum", world, dev)  # busy steps are untimed
    return {"metric": "env-steps/sec", "value": counted / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "dtype": "f64 env / bf16 policy", "clocks": clocks,
            "config": {"workload": "cfg4: PPO rollout, state norm + transformer policy forward + masked discrete sampling + RS plan hand-off + full env step",
                       "envs_per_gpu": env.n, "policy": what, "modalities": 4 if use_img else 3, "glue": eng.glue,
                       "policy_params": sum(p.numel() for p in actor.parameters())}}


def measure_sac(env, rank, local_rank, world, dev, steps, warmup):
    """BASELINE cfg 5: SAC-style acting + device replay + one update every 8 env steps; the gradients of actor + twin critics +
    temperature travel in ONE flat NCCL all-reduce per update that overlaps the following rollout steps."""
    import torch
    from hope_b200 import learner
    loop = learner.SacRollout(env, world=world, seed=0)
    loop.run(max(warmup, 2 * loop.update_every))
    torch.cuda.synchronize()
    c0, u0 = env.counters(), loop.updates
    ran = {"n": 0}

    def run():
        loop.run(steps); ran["n"] += steps

    def busy():
        loop.run(loop.update_every); ran["n"] += loop.update_every

    ms, clocks = timed_with_clocks(rank, local_rank, world, dev, run, busy, loop.update_every / steps)
    c1 = env.counters()
    counted = reduce_scalar((c1["env_steps"] - c0["env_steps"]) * steps / max(1, ran["n"]), "sum", world, dev)
    w = torch.cat([p.detach().reshape(-1).float() for p in loop.learner.actor.parameters()])
    wsum = float(w.double().sum())
    spread = reduce_scalar(wsum, "max", world, dev) + reduce_scalar(-wsum, "max", world, dev)
    ar_ms = reduce_scalar(loop.learner.time_allreduce(5), "max", world, dev)
    return {"metric": "env-steps/sec", "value": counted / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "dtype": "f64 env / bf16 nets", "clocks": clocks,
            "allreduce_ms_per_update": ar_ms, "allreduce_floats": loop.learner.reducer.numel,
            "config": {"workload": "cfg5: SAC rollout + device replay + update every 8 env steps (batch 8192/GPU), one flat NCCL gradient all-reduce per update",
                       "envs_per_gpu": env.n, "updates_in_region": (loop.updates - u0) * steps // max(1, ran["n"]), "replica_weight_spread": spread,
                       "allreduce": "sum over ranks of the actor + 2 critics + log_alpha gradients, fp32, asynchronous (overlaps the next 8 rollout steps); "
                                    "allreduce_ms_per_update is the same bucket reduced alone, not overlapped"}}


def _device_loop(env, actions, lo, hi, step_fn=None):
    step = step_fn or env.step
    for k in range(lo, hi):
        step(actions[k % len(actions)])


def run_single_config(args, rank, local_rank, world, dev):
    """--config rollout | sac | dlp | cfg2 | image: one line for that configuration, with its clock record"""
    import torch
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    K, W = args.steps, max(3, args.warmup)
    cfg = args.config
    if cfg in ("rollout", "sac"):
        n = args.envs
        env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", scene_seed(rank)), device=local_rank, auto_reset=True,
                                use_img_observation=(cfg == "rollout" and args.image))
        line = measure_rollout(env, rank, local_rank, world, dev, K, W, use_img=args.image) if cfg == "rollout" else \
            measure_sac(env, rank, local_rank, world, dev, K, W)
    else:
        if cfg == "dlp":
            from hope_b200 import dlp
            n = args.envs if args.envs != ENVS_PER_GPU else 16384
            cases = dlp.cases_from_fixture(np.load(os.path.join(ROOT, "tests", "golden", "dlp_cases.npz")))
            sc = dlp.prepare_scenes(cases, np.arange(2 * n) % len(cases), seed=scene_seed(rank))
            env = BatchedParkingEnv(n, scenes=sc, device=local_rank, auto_reset=True)
            nobs = (sc["nverts"] > 0).sum(axis=1)
            workload = {"workload": "row f3: full step on Dragon Lake Parking scenes, 128-ring build", "envs_per_gpu": n, "data": "dlp fixture cases",
                        "obstacle_rings_per_scene": {"min": int(nobs.min()), "mean": float(nobs.mean()), "max": int(nobs.max())}}
            step_fn = None
        elif cfg == "cfg2":
            n = args.envs if args.envs != ENVS_PER_GPU else 4096
            env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", scene_seed(rank)), device=local_rank, auto_reset=True)
            workload = {"workload": "cfg2: kinematics + ring collision (+ arrival) only, level Normal", "envs_per_gpu": n}
            step_fn = env.step_kinematics_collision
        else:  # image
            n = args.envs
            env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", scene_seed(rank)), device=local_rank, auto_reset=True,
                                    use_img_observation=True)
            workload = {"workload": "row f1: cfg-3 full step + image observation (3x64x64 uint8 per env)", "envs_per_gpu": n,
                        "l2": "image output alone is 805 MB per step at 65 536 envs, larger than L2"}
            step_fn = None
        env.reset()
        gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
        actions = torch.rand((K + W, n, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        _device_loop(env, actions, 0, W, step_fn)
        torch.cuda.synchronize()
        c0 = env.counters()
        env.profile(True); env.profile_read()
        extra = {"n": 0}

        def busy():
            _device_loop(env, actions, 0, 4, step_fn); extra["n"] += 4

        ms, clocks = timed_with_clocks(rank, local_rank, world, dev, lambda: _device_loop(env, actions, W, W + K, step_fn), busy, 4.0 / K)
        prof = env.profile_read(); env.profile(False)
        c1 = env.counters()
        steps = reduce_scalar((c1["env_steps"] - c0["env_steps"]) * K / (K + extra["n"]), "sum", world, dev)
        kms = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
        line = {"metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "dtype": "f64" if cfg != "image" else "f64 step, u8 image", "clocks": clocks,
                "gpu_launches": (c1["kernel_launches"] - c0["kernel_launches"]) * K // (K + extra["n"]), "kernels_ms_per_launch": kms, "config": workload}
        if cfg in ("cfg2", "image"):
            kern = "k_advance" if cfg == "cfg2" else "k_render"
            peak, src = hbm_peak()
            if kms.get(kern):
                ach = ALGO_BYTES[kern] * n / (kms[kern] * 1e-3) / 1e9
                line["roofline"] = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                    "algorithmic_bytes_per_env_step": ALGO_BYTES[kern], "peak_source": src}
    if rank == 0:
        line.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": line["config"].pop("data", "synthetic")})
        print(json.dumps(line))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="scenes per GPU (default: the BASELINE cfg-3 size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg-4 / cfg-5 blocks of the main line")
    ap.add_argument("--device-only", action="store_true", help="profiling runs: stop after the device-resident timing (no e2e, no CPU baseline)")
    ap.add_argument("--image", action="store_true", help="--config rollout: 4-modal policy with the image observation on")
    ap.add_argument("--config", default="step", choices=["step", "rollout", "sac", "dlp", "cfg2", "image"],
                    help="step: BASELINE cfg 3 (default, the headline metric); rollout: cfg 4; sac: cfg 5; dlp / cfg2 / image: rows f3, cfg 2, f1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hope_b200 import capi
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes

    rank, local_rank, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("HOPE_B200_HOST_THREADS", str(host_threads_for_rank(world)))  # mask expansion threads: the box's CPUs shared by the ranks
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, K, W = args.envs, args.steps, max(3, args.warmup)
    dev = torch.device("cuda", local_rank)
    if args.config != "step":
        run_single_config(args, rank, local_rank, world, dev)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # scene id -> GPU: rank r owns scenes [r*2n, (r+1)*2n) of the global synthetic pool
    scenes = generate_scenes(2 * n, "mix", scene_seed(rank))
    env = BatchedParkingEnv(n, scenes=scenes, device=local_rank, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    actions = torch.rand((K + W, n, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return reduce_scalar(x, "max", world, dev)

    def sum_over_ranks(x):
        return reduce_scalar(x, "sum", world, dev)

    # ---- device-resident timing ----------------------------------------------------------------
    _device_loop(env, actions, 0, W)
    barrier()
    c0 = env.counters()
    extra = {"n": 0}

    def busy():
        _device_loop(env, actions, 0, 8); extra["n"] += 8

    ms, clocks = timed_with_clocks(rank, local_rank, world, dev, lambda: _device_loop(env, actions, W, W + K), busy, 8.0 / K)
    c1 = env.counters()
    ran = K + extra["n"]
    steps_local = (c1["env_steps"] - c0["env_steps"]) * K / ran  # env-steps with an action (auto-reset steps excluded); busy steps are untimed
    assert 0 < steps_local <= n * K, (steps_local, n, K)
    launches = (c1["kernel_launches"] - c0["kernel_launches"]) * K // ran
    total_steps = sum_over_ranks(float(steps_local))
    value = total_steps / (ms * 1e-3)

    # ---- per-kernel durations: a second pass over the same kind of steps, every kernel alone on one stream -------------------------
    env.profile(True, serial=True); env.profile_read()
    _device_loop(env, actions, W, W + min(K, 20))
    serial = env.profile_read()
    env.profile(True, serial=False); env.profile_read()
    _device_loop(env, actions, W, W + min(K, 20))
    live = env.profile_read()
    env.profile(False)
    if args.device_only:
        if rank == 0:
            print(json.dumps({"device_only": True, "value": value, "ms_per_step": ms / K,
                              "kernels_ms_serial": {k: (v[0] / v[1] if v[1] else None) for k, v in serial.items()}}))
        env.close()
        return

    # ---- end-to-end through the host-buffer C ABI ---------------------------------------------------
    h_actions = actions[:, :, :].cpu().numpy()
    for k in range(W):
        env.step_host(h_actions[k])
    barrier()
    c2 = env.counters()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        env.step_host(h_actions[k])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    c3 = env.counters()
    e2e_steps = sum_over_ranks(float(c3["env_steps"] - c2["env_steps"]))
    wire = env.host_wire_info()  # counted by the library from the copies of the last step (the kept lidar values are data dependent)
    h2d, d2h = wire["h2d_bytes"], wire["d2h_bytes"]
    # context for the e2e number: what a plain pinned D2H copy of the lidar buffer achieves on this box while EVERY rank copies
    pin = torch.empty_like(env.out["lidar"], device="cpu").pin_memory()
    barrier()
    t1 = time.perf_counter()
    for _ in range(5):
        pin.copy_(env.out["lidar"], non_blocking=True)
    torch.cuda.synchronize()
    pcie_local = 5 * pin.numel() * 8 / (time.perf_counter() - t1) / 1e9
    pcie_min, pcie_sum = -max_over_ranks(-pcie_local), sum_over_ranks(pcie_local)

    secondary = None
    if not args.no_secondary:
        K2, W2 = min(K, 16), 4
        secondary = {"cfg4": measure_rollout(env, rank, local_rank, world, dev, K2, W2),
                     "cfg5": measure_sac(env, rank, local_rank, world, dev, 64, 16)}  # 8 updates in the timed region (0.2 s): one host hiccup in 16 steps had tripled the figure
        # cfg 4 as the reference ships it (USE_IMG: 4-modal actor, image observation rendered every step): the same scenes, a
        # second env with the image stage on
        env_img = BatchedParkingEnv(n, scenes=scenes, device=local_rank, auto_reset=True, use_img_observation=True)
        secondary["cfg4_img"] = measure_rollout(env_img, rank, local_rank, world, dev, K2, 22, use_img=True)  # 22 warm-up steps: the trajectory boxes k_render draws are all there
        secondary["cfg4_img"]["config"]["workload"] = "cfg4 with USE_IMG: k_render + image conv stack kernel + 4-modal actor kernel in the loop"
        env_img.close()
        del env_img

    if rank == 0:
        ser = {k: (v[0] / v[1] if v[1] else None) for k, v in serial.items()}
        ser_sum = sum(v for v in ser.values() if v)
        dom = max((k for k in ser if ser[k]), key=lambda k: ser[k])
        peak, peak_src = hbm_peak()
        achieved = ALGO_BYTES[dom] * n / (ser[dom] * 1e-3) / 1e9
        import ctypes
        fp64 = ctypes.c_double(0.0)
        capi.check(capi.load_library().hope_fp64_peak_tflops(local_rank, ctypes.byref(fp64)))
        counts = {}
        cpath = os.path.join(ROOT, "profiles", "ncu_counts.json")
        if os.path.exists(cpath):
            counts = json.load(open(cpath))
        flops_step = counts.get("fp64_flops_per_step_65536")
        fp64_block = None
        if flops_step:
            tfl = flops_step * (n / 65536.0) / (ms / K * 1e-3) / 1e12
            fp64_block = {"achieved_tflops": tfl, "peak_tflops_measured": fp64.value, "frac": tfl / fp64.value if fp64.value else None,
                          "flops_per_env_step": flops_step / 65536.0,
                          "source": f"executed DADD + DMUL + 2 DFMA thread instructions of all kernels of one step, ncu at commit {counts.get('commit')} "
                                    "(profiles/ncu_counts.json), over this run's live ms_per_step; the peak is an 8-chain DFMA loop measured in this run"}
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "scene_pool_per_gpu": 2 * n, "parallelism": f"scenes sharded x{world}, no data-path collective",
                       "l2": "per-step working set (scene pool 2N x 1.7 KB + outputs N x 1.5 KB = 320 MB) exceeds the 126 MB L2; no explicit flush",
                       "counted": "env-steps with an action; auto-reset steps excluded"},
            "e2e": {"value": e2e_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "hope_step_host (pinned host buffers, synchronous; lossless narrow wire format: the float64 mask crosses PCIe as uint8 step counts, "
                           "the float64 lidar as 120 flag bits + the beams that differ from the per-ray no-hit constant; "
                           f"{wire['host_threads']} host threads rebuild both arrays inside the call, AVX-512 {'on' if wire['avx512'] else 'off'})",
                    "wire": wire,
                    "host_bytes_delivered_per_step": int(env.n * 1436), "ms_per_step": 1e3 * e2e_s / K,
                    "d2h_gbs_per_rank_in_step": d2h / (e2e_s / K) / 1e9,
                    "plain_d2h_copy_gbs": {"per_rank_min": pcie_min, "all_ranks_sum": pcie_sum, "note": "all ranks copying at the same time"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "kernels_ms_per_launch": ser,
            "kernels_ms_overlapped_in_live_step": {k: (v[0] / v[1] if v[1] else None) for k, v in live.items()},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (counts.get("dram_bytes_per_launch") or {}).get(dom),
                         "traffic_source": (f"ncu --set full dram__bytes_read+write of {dom} at commit {counts.get('commit')} (profiles/ncu_counts.json); "
                                            "dram__bytes understates a kernel whose outputs are still in the 126 MB L2 when it ends") if counts else None,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES[dom], "peak_source": peak_src,
                         "kernel_share_of_step_serialised": ser[dom] / ser_sum,
                         "full_step": {"achieved": FULL_STEP_BYTES * n / (ms / K * 1e-3) / 1e9, "frac": FULL_STEP_BYTES * n / (ms / K * 1e-3) / 1e9 / peak,
                                       "algorithmic_bytes_per_env_step": FULL_STEP_BYTES},
                         "note": "float64 ALU/issue-bound path: the HBM fraction is small by construction (SURVEY.md §8d); the binding resource is in `fp64`",
                         "timing": "CUDA events around each launch of a serialised pass (hope_profile_enable 2), live loop, warm L2"},
            "fp64": fp64_block,
        }
        if secondary:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, counted, dt = cpu_oracle_rate(n, CPU_SAMPLE_STEPS, threads)
            one_rate, one_counted, one_dt = cpu_oracle_rate(ONE_CORE_ENVS, ONE_CORE_STEPS, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"C oracle, OpenMP x{threads}, {n} scenes x {CPU_SAMPLE_STEPS} steps = {counted} env-steps in {dt:.1f} s "
                                              "(288 reference-generated scenes tiled, tests/golden/scenes_*.npz)",
                                    "one_core": {"value": one_rate, "unit": "env-steps/s", "cores": 1,
                                                 "sample": f"1 thread, {ONE_CORE_ENVS} scenes x {ONE_CORE_STEPS} steps = {one_counted} env-steps in {one_dt:.1f} s"},
                                    "python_reference": PY_REFERENCE_NOTE}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
