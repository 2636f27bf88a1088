#!/usr/bin/env python
"""hope_step_host timing under different wire-format / pipelining settings, one fresh process per setting (the settings
are read at hope_create).  Wall clock around K synchronous host steps after W warm-up steps, cfg-3 workload.

Usage (GPU box):  python profiles/tools/e2e_sweep.py --out gpurun_out/e2e_sweep.jsonl "HOPE_B200_HOST_CHUNKS=4" "HOPE_B200_HOST_LIDAR_PACK=0" ...
                  (each argument: space-separated VAR=VALUE pairs of one setting; "" = defaults)
Not a bench value (no clock sampling); the numbers rank the settings against each other.
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))


def child(envs, steps, warmup):
    sys.path.insert(0, ROOT)
    import torch
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    scenes = generate_scenes(2 * envs, "mix", 42)
    env = BatchedParkingEnv(envs, scenes=scenes, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(1234)
    act = (torch.rand((steps + warmup, envs, 2), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1).cpu().numpy()
    for k in range(warmup):
        env.step_host(act[k])
    torch.cuda.synchronize()
    per = []
    t0 = time.perf_counter()
    for k in range(warmup, warmup + steps):
        t1 = time.perf_counter()
        env.step_host(act[k])
        per.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    in_order = [round(1e3 * x, 3) for x in per[:12]]
    per.sort()
    w = env.host_wire_info()
    first = [round(1e3 * x, 3) for x in per[:0]]
    print(json.dumps({"ms_per_step": 1e3 * dt / steps, "ms_median": 1e3 * per[len(per) // 2], "ms_p10": 1e3 * per[len(per) // 10],
                      "first_steps_ms": in_order, "env_steps_per_s_nominal": envs * steps / dt, "wire": w, "d2h_gbs": w["d2h_bytes"] * steps / dt / 1e9}))
    env.close()


def hostinfo():
    """what kind of host this is: the e2e number moves with the host's store bandwidth (the expansion writes 85 MB per step)"""
    sys.path.insert(0, ROOT)
    import numpy as np
    from hope_b200 import capi, tables
    lib = capi.load_library()
    n = 65536
    rng = np.random.default_rng(0)
    nohit = 10.0 - tables.host_tables()["lidar_base"]
    keep = rng.random((n, 120)) < 0.57
    bits = np.zeros((n, 4), dtype=np.uint32)
    for j in range(120):
        bits[:, j // 32] |= keep[:, j].astype(np.uint32) << np.uint32(j % 32)
    cnt = keep.sum(1)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint32)
    packed = rng.random(int(cnt.sum()) + 8)
    out = np.zeros((n, 120))
    best = 1e9
    for _ in range(5):
        t = time.perf_counter()
        lib.hope_expand_lidar(bits.ctypes.data, off.ctypes.data, packed.ctypes.data, nohit.ctypes.data, out.ctypes.data, n, 0)
        best = min(best, time.perf_counter() - t)
    a = np.zeros(32 << 20); b = np.zeros(32 << 20)
    cp = 1e9
    for _ in range(3):
        t = time.perf_counter(); np.copyto(b, a); cp = min(cp, time.perf_counter() - t)
    model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
    mhz = [float(l.split(":", 1)[1]) for l in open("/proc/cpuinfo") if l.startswith("cpu MHz")]
    print(json.dumps({"hostinfo": {"cpus": len(model), "model": model[0] if model else None, "mhz_max": max(mhz) if mhz else None,
                                   "expand_lidar_65536_one_thread_ms": 1e3 * best, "numpy_copy_256MB_gbs": 2 * a.nbytes / cp / 1e9}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("settings", nargs="*")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args.envs, args.steps, args.warmup)
    hostinfo()
    for setting in (args.settings or [""]):
        env = dict(os.environ)
        for kv in setting.split():
            k, v = kv.split("=", 1)
            env[k] = v
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--envs", str(args.envs), "--steps", str(args.steps),
                            "--warmup", str(args.warmup)], env=env, capture_output=True, text=True)
        try:
            res = json.loads(p.stdout.strip().splitlines()[-1])
        except (IndexError, ValueError):
            res = {"error": (p.stderr or p.stdout)[-600:]}
        traces = [json.loads(l) for l in p.stderr.splitlines() if l.startswith('{"host_step"')]  # HOPE_B200_HOST_TRACE=k
        if traces:
            res["trace"] = traces[-1]
        line = json.dumps({"env": setting, "result": res})
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")


if __name__ == "__main__":
    main()
