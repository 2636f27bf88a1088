#!/usr/bin/env python
"""bench.py — env-steps/s of the batched ParkingEnv step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

Workload = BASELINE.json configs[2]: 65 536 parallel scenes per GPU (levels Normal/Complex/Extrem
cycled), FULL step: kinematics + ring collision + arrival + status/reward, 120-beam LiDAR raycast,
42-action mask sweep, Reeds-Shepp search; float64 random actions U(-1,1)^2; finished envs take
their next scene from a pre-generated pool of 2N scenes on the following step (those reset
steps are NOT counted as env-steps).  One "step" = one pass over all scenes of the rank.

Timed regions
  value  device-resident: actions already in HBM, hope_step on the current stream, CUDA events,
         max over ranks.
  e2e    hope_step_host: pinned host actions -> H2D, step, D2H of the observation/reward/done/RS
         buffers, synchronised; wall clock around the synchronous calls, max over ranks.
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


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"
CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS = 4096, 128  # ~12 s of CPU work on the GPU box (16 host threads)


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


def pin_to_gpu_numa(local_rank):
    """Bind this rank's host threads (and therefore its pinned staging buffers, first-touch) to the CPUs
    that are local to its GPU's PCIe root: with 8 ranks the D2H streams otherwise cross sockets."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


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

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

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


def cpu_oracle_rate(n_envs, steps, nthreads, seed=42):
    """env-steps/s of the C oracle (oracle/c/parking_oracle.c, OpenMP over scenes) on a bounded
    sample of the same workload; reset steps excluded like on the GPU arm."""
    from hope_b200.batched_env import generate_scenes
    from oracle import parking_oracle as po
    sc = generate_scenes(n_envs, "mix", seed)
    env = po.OracleEnv(sc["start"], sc["dest"], sc["bounds"], sc["obs"], sc["nverts"], nthreads=nthreads)
    env.reset_step()
    rng = np.random.default_rng(seed)
    env.step(rng.uniform(-1, 1, size=(n_envs, 2)))  # warm-up (page in the 4 MB table)
    done = np.zeros(n_envs, dtype=bool)
    counted, t0 = 0, time.perf_counter()
    for _ in range(steps):
        act = rng.uniform(-1, 1, size=(n_envs, 2))
        if done.any():  # next-step auto-reset, same convention as the CUDA path
            env.reset_state(np.where(done)[0])
        out = env.step(act, has_action=(~done).astype(np.uint8))
        counted += int((~done).sum())
        done = out["status"] != 1
    dt = time.perf_counter() - t0
    return counted / dt, counted, dt


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(max(0, args.warmup - 1)):
        cpu_oracle_rate(CPU_SAMPLE_ENVS, 1, threads)
    rate, counted, dt = cpu_oracle_rate(CPU_SAMPLE_ENVS, max(1, args.steps), threads)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{CPU_SAMPLE_ENVS} scenes x {args.steps} steps per run"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"C oracle (oracle/c/parking_oracle.c, OpenMP x{threads}) on {CPU_SAMPLE_ENVS} scenes x {args.steps} steps "
                                   f"= {counted} env-steps; the reference itself is single-process Python at 65-87 env-steps/s/core (BASELINE.md §2)"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_rollout(args, rank, local_rank, world, dev):
    """BASELINE cfg 4: PPO acting loop, 65 536 envs x K steps: state norm -> transformer policy (bf16 autocast,
    tensor-core GEMMs through stock PyTorch) -> masked discrete sampling -> RS plan hand-off -> env step."""
    import torch
    from hope_b200 import rollout
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n, K, W = args.envs, args.steps, max(3, args.warmup)
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", scene_seed(rank)), device=local_rank, auto_reset=True)
    actor = rollout.ReferenceShapedActor().to(dev)
    eng = rollout.RolloutEngine(env, actor, seed=rank)
    eng.collect(W)
    torch.cuda.synchronize()
    c0 = env.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.collect(K)
    e1.record()
    torch.cuda.synchronize()
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    c1 = env.counters()
    steps = reduce_scalar(float(c1["env_steps"] - c0["env_steps"]), "sum", world, dev)
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 env / bf16 policy",
            "data": "synthetic", "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "config": {"workload": "cfg4: PPO rollout, 65536 envs/GPU, transformer policy forward + masked sampling + RS plan hand-off + full env step",
                       "envs_per_gpu": n, "policy": "ReferenceShapedActor (MultiObsEmbedding shapes, 3 modalities), random init",
                       "policy_params": sum(p.numel() for p in actor.parameters())}}))
    env.close()


def run_dlp(args, rank, local_rank, world, dev):
    """Row f3: full step on Dragon Lake Parking scenes (31-119 obstacle rings, 128-ring build of the library).
    Scenes: the 16 fixture cases of tests/golden/dlp_cases.npz, each prepared with different random starts."""
    import torch
    from hope_b200 import dlp
    from hope_b200.batched_env import BatchedParkingEnv
    n = args.envs if args.envs != ENVS_PER_GPU else 16384
    K, W = args.steps, max(3, args.warmup)
    cases = dlp.cases_from_fixture(np.load(os.path.join(ROOT, "tests", "golden", "dlp_cases.npz")))
    sc = dlp.prepare_scenes(cases, np.arange(2 * n) % len(cases), seed=scene_seed(rank))
    env = BatchedParkingEnv(n, scenes=sc, device=local_rank, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
    actions = torch.rand((K + W, n, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    for k in range(W):
        env.step(actions[k])
    torch.cuda.synchronize()
    c0 = env.counters()
    env.profile(True); env.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W, W + K):
        env.step(actions[k])
    e1.record()
    torch.cuda.synchronize()
    prof = env.profile_read()
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    c1 = env.counters()
    steps = reduce_scalar(float(c1["env_steps"] - c0["env_steps"]), "sum", world, dev)
    if rank == 0:
        nobs = (sc["nverts"] > 0).sum(axis=1)
        print(json.dumps({
            "metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "dlp fixture cases",
            "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "kernels_ms_per_launch": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()},
            "config": {"workload": "row f3: full step on Dragon Lake Parking scenes, 128-ring build", "envs_per_gpu": n,
                       "obstacle_rings_per_scene": {"min": int(nobs.min()), "mean": float(nobs.mean()), "max": int(nobs.max())}}}))
    env.close()


def run_cfg2(args, rank, local_rank, world, dev):
    """BASELINE cfg 2: 4 096 parallel scenes (level Normal), kinematics + ring collision (+ arrival) only."""
    import torch
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n = args.envs if args.envs != ENVS_PER_GPU else 4096
    K, W = args.steps, max(3, args.warmup)
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", scene_seed(rank)), device=local_rank, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(3 + rank)
    actions = torch.rand((K + W, n, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    for k in range(W):
        env.step_kinematics_collision(actions[k])
    torch.cuda.synchronize()
    c0 = env.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W, W + K):
        env.step_kinematics_collision(actions[k])
    e1.record()
    torch.cuda.synchronize()
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    c1 = env.counters()
    steps = reduce_scalar(float(c1["env_steps"] - c0["env_steps"]), "sum", world, dev)
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "roofline": {"bound": "hbm", "kernel": "k_advance", "achieved": ALGO_BYTES["k_advance"] * n / (ms / K * 1e-3) / 1e9, "unit": "GB/s",
                         "note": "launch-latency bound at 4 096 scenes (one 32-block kernel per step)"},
            "config": {"workload": "cfg2: kinematics + ring collision (+ arrival) only, level Normal", "envs_per_gpu": n}}))
    env.close()


def run_image(args, rank, local_rank, world, dev):
    """Row f1: the cfg-3 full step plus the ego-centric image observation (k_render), device-resident."""
    import torch
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n = args.envs
    K, W = args.steps, max(3, args.warmup)
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", scene_seed(rank)), device=local_rank, auto_reset=True,
                            use_img_observation=True)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    actions = torch.rand((K + W, n, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    for k in range(W):
        env.step(actions[k])
    torch.cuda.synchronize()
    c0 = env.counters()
    env.profile(True); env.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W, W + K):
        env.step(actions[k])
    e1.record()
    torch.cuda.synchronize()
    prof = env.profile_read()
    env.profile(False)
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    c1 = env.counters()
    steps = reduce_scalar(float(c1["env_steps"] - c0["env_steps"]), "sum", world, dev)
    if rank == 0:
        kms = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
        peak, src = hbm_peak()
        ach = ALGO_BYTES["k_render"] * n / (kms["k_render"] * 1e-3) / 1e9
        print(json.dumps({
            "metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 step, u8 image",
            "data": "synthetic", "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"], "kernels_ms_per_launch": kms,
            "roofline": {"bound": "hbm", "kernel": "k_render", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "algorithmic_bytes_per_env_step": ALGO_BYTES["k_render"], "peak_source": src,
                         "note": "shared-memory / issue bound: 16 384 screen samples per env resolved on chip, 12 KB written"},
            "config": {"workload": "row f1: cfg-3 full step + image observation (3x64x64 uint8 per env)", "envs_per_gpu": n,
                       "l2": "image output alone is 805 MB per step at 65 536 envs, larger than L2"}}))
    env.close()


def run_sac(args, rank, local_rank, world, dev):
    """BASELINE cfg 5: 65 536 envs per GPU, SAC-style acting + replay + one update every 8 env steps, gradients of
    actor + twin critics reduced with a single NCCL all-reduce that overlaps the following rollout steps."""
    import torch
    from hope_b200 import learner
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n, K, W = args.envs, args.steps, max(3, args.warmup)
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", scene_seed(rank)), device=local_rank, auto_reset=True)
    loop = learner.SacRollout(env, world=world, seed=0)
    loop.run(max(W, 2 * loop.update_every))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    c0, u0 = env.counters(), loop.updates
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(K)
    e1.record()
    torch.cuda.synchronize()
    ms = reduce_scalar(e0.elapsed_time(e1), "max", world, dev)
    c1 = env.counters()
    steps = reduce_scalar(float(c1["env_steps"] - c0["env_steps"]), "sum", world, dev)
    # replicas must stay identical: same init, same averaged gradients
    w = torch.cat([p.detach().reshape(-1).float() for p in loop.learner.actor.parameters()])
    spread = reduce_scalar(float(w.double().sum()), "max", world, dev) - (-reduce_scalar(-float(w.double().sum()), "max", world, dev))
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec", "value": steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 env / bf16 nets",
            "data": "synthetic", "gpu_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "config": {"workload": "cfg5: SAC rollout + device replay + update every 8 env steps (batch 8192/GPU), one flat NCCL gradient all-reduce per update",
                       "envs_per_gpu": n, "updates": loop.updates - u0, "allreduce_floats": loop.learner.reducer.numel,
                       "replica_weight_spread": spread}}))
    env.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="scenes per GPU (default: the BASELINE cfg-3 size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="profiling runs: stop after the device-resident timing (no e2e, no CPU baseline, no JSON line)")
    ap.add_argument("--config", default="step", choices=["step", "rollout", "sac", "dlp", "cfg2", "image"],
                    help="step: BASELINE cfg 3 (default, the headline metric); rollout: cfg 4, PPO acting loop with the transformer policy")
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
    numa = pin_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, K, W = args.envs, args.steps, max(3, args.warmup)
    dev = torch.device("cuda", local_rank)
    if args.config == "rollout":
        return run_rollout(args, rank, local_rank, world, dev)
    if args.config == "sac":
        return run_sac(args, rank, local_rank, world, dev)
    if args.config == "dlp":
        return run_dlp(args, rank, local_rank, world, dev)
    if args.config == "cfg2":
        return run_cfg2(args, rank, local_rank, world, dev)
    if args.config == "image":
        return run_image(args, rank, local_rank, world, dev)

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
    for k in range(W):
        env.step(actions[k])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 3.0:  # let nvidia-smi deliver its first line,
            env.step(actions[0])                                  # keeping the GPU busy (untimed, uncounted)
        sampler.lines.clear()
    barrier()
    c0 = env.counters()
    env.profile(True)  # per-kernel CUDA events inside the timed region
    env.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(W, W + K):
        env.step(actions[k])
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = env.profile_read()
    env.profile(False)
    c1 = env.counters()
    steps_local = c1["env_steps"] - c0["env_steps"]  # env-steps with an action (auto-reset steps excluded)
    assert 0 < steps_local <= n * K, (steps_local, n, K)
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    ms = max_over_ranks(ms_local)
    total_steps = sum_over_ranks(float(steps_local))
    value = total_steps / (ms * 1e-3)

    if args.device_only:
        if rank == 0:
            print(json.dumps({"device_only": True, "value": value, "ms_per_step": ms / K}))
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
    h2d, d2h = env.host_io_bytes()
    # context for the e2e number: what a plain pinned D2H copy of the lidar buffer achieves on this box
    pin = torch.empty_like(env.out["lidar"], device="cpu").pin_memory()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(5):
        pin.copy_(env.out["lidar"], non_blocking=True)
    torch.cuda.synchronize()
    pcie_gbs = 5 * pin.numel() * 8 / (time.perf_counter() - t1) / 1e9

    if rank == 0:
        dom = max(prof, key=lambda k: prof[k][0])
        dom_ms, dom_launches = prof[dom]
        peak, peak_src = hbm_peak()
        achieved = ALGO_BYTES[dom] * n / (dom_ms / max(1, dom_launches) * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        import ctypes
        fp64 = ctypes.c_double(0.0)
        capi.check(capi.load_library().hope_fp64_peak_tflops(local_rank, ctypes.byref(fp64)))
        pipe = None
        ppath = os.path.join(ROOT, "profiles", "fp64_pipe_pct.json")
        if os.path.exists(ppath):
            pipe = json.load(open(ppath)).get(dom)
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "scene_pool_per_gpu": 2 * n, "parallelism": f"scenes sharded x{world}, no data-path collective",
                       "l2": "per-step working set (scene pool 2N x 1.7 KB + outputs N x 1.5 KB = 320 MB) exceeds the 126 MB L2; no explicit flush",
                       "counted": "env-steps with an action; auto-reset steps excluded"},
            "e2e": {"value": e2e_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "hope_step_host (pinned host buffers, synchronous; the float64 mask crosses PCIe as uint8 step counts and is expanded by 6 host threads inside the call)", "host_bytes_delivered_per_step": int(env.n * 1436), "ms_per_step": 1e3 * e2e_s / K,
                    "plain_d2h_copy_gbs": pcie_gbs,
                    "host_cpus_rank0": numa},
            "gpu_launches": launches,
            "clocks": clocks,
            "kernels_ms_per_launch": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_env_step": ALGO_BYTES[dom], "peak_source": peak_src,
                         "note": "float64 ALU/latency-bound path: HBM fraction is small by construction (SURVEY.md §8d)",
                         "fp64_fma_peak_tflops_measured": fp64.value, "fp64_pipe_active_pct_ncu": pipe},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, counted, dt = cpu_oracle_rate(CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"C oracle, OpenMP x{threads}, {CPU_SAMPLE_ENVS} scenes x {CPU_SAMPLE_STEPS} steps = {counted} env-steps in {dt:.1f} s; "
                                              "reference Python env: 65-87 env-steps/s/core (BASELINE.md §2, survey probe)"}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
