#!/usr/bin/env python
"""Work counters of the BASELINE cfg-3 step from an instrumented build of the kernels (-DHOPE_STATS).

The default library contains none of the counting code (its SASS is identical with and without the #ifdef blocks).
This tool builds hope_b200/csrc with -DHOPE_STATS into profiles/tools/_stats/, steps the same workload as bench.py
(65 536 mixed-level scenes, auto-reset from a 2N pool, uniform random actions) and prints one JSON object:
where k_rs_check's words end (exit round, bounds or obstacle, obstacle index), how full its warps are, and how many
(quadrant, edge) iterations / screened rays / table fetches k_observe does per env.  Numbers are per-step work, not
times; nothing printed here is a bench value.

Usage (GPU box):  python profiles/tools/kernel_stats.py [--envs 65536] [--warmup 10] [--steps 30] [--out FILE]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
STATS_LIB = os.path.join(HERE, "_stats", "libhope_b200_stats.so")


def build_stats_lib(force=False):
    from hope_b200 import build as hb
    srcs = [os.path.join(hb.CSRC, s) for s in hb.SOURCES]
    deps = [os.path.join(hb.CSRC, d) for d in hb.DEPS]
    if not force and os.path.exists(STATS_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(STATS_LIB) for d in deps):
        return STATS_LIB
    os.makedirs(os.path.dirname(STATS_LIB), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([hb.nvcc_path()] + hb.NVCC_FLAGS + ["-DHOPE_STATS", "-o", STATS_LIB] + srcs, env=env)
    return STATS_LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    lib_path = build_stats_lib()
    if args.build_only:
        print(lib_path)
        return
    import torch
    from hope_b200 import build as hb, capi
    hb.VARIANTS[16] = lib_path  # this process steps the instrumented build
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    lib = capi.load_library(16)
    lib.hope_debug_stats.restype = C.c_int
    lib.hope_debug_stats.argtypes = [C.POINTER(C.c_uint64 * 64), C.c_int]
    n = args.envs
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", 42), auto_reset=True)
    env.reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(1234)
    act = torch.rand((args.warmup + args.steps, n, 2), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1
    for k in range(args.warmup):
        env.step(act[k])
    buf = (C.c_uint64 * 64)()
    capi.check(lib.hope_debug_stats(C.byref(buf), 1))
    c0 = env.counters()
    for k in range(args.warmup, args.warmup + args.steps):
        env.step(act[k])
    capi.check(lib.hope_debug_stats(C.byref(buf), 0))
    c1 = env.counters()
    s = [int(v) for v in buf]
    K = args.steps
    words, bad = s[0], s[1]
    rounds = s[13]
    exit_hist = s[2:10]
    bad_rounds = sum((r + 1) * c for r, c in enumerate(exit_hist))  # rounds spent on words that end with a hit (first chunk)
    rec = {
        "workload": f"cfg3, {n} envs, {K} steps after {args.warmup} warm-up steps",
        "env_steps": c1["env_steps"] - c0["env_steps"],
        "k_rs_check": {
            "words_per_step": words / K, "bad_fraction": bad / max(1, words),
            "mean_samples_of_words_walked_to_their_end": s[11] / max(1, s[35]), "words_walked_to_their_end": s[35] / max(1, words),
            "words_longer_than_one_chunk": s[10],
            "rounds_per_word": rounds / max(1, words),
            "valid_lanes_per_round": s[12] / max(1, rounds),
            "exit_round_histogram_of_bad_words": exit_hist,
            "share_of_rounds_spent_on_bad_words": bad_rounds / max(1, rounds),
            "exits_by_bounds": s[14], "exits_by_obstacle": s[15],
            "exit_obstacle_index_histogram": s[16:32],
            "exit_obstacle_edge_histogram": s[36:40],  # per-edge vote (HOPE_CHK_EDGE_EXIT=1): after which edge of the obstacle
            "obstacle_iterations_per_round": s[32] / max(1, rounds),
            "obstacle_iterations_with_edge_work": s[34] / max(1, s[32]),
            "lanes_in_edge_loop_when_any": s[33] / max(1, s[34]),
        },
        "k_observe": {
            "envs": s[40] / K,
            "edges_per_env": s[42] / max(1, s[40]),
            "quadrant_edge_iterations_per_env": s[41] / max(1, s[40]),
            "active_beams_per_env": s[43] / max(1, s[40]),
            "screen2_passes_per_env": s[46] / max(1, s[40]),
            "screened_rays_per_env": s[44] / max(1, s[40]),
            "table_fetches_per_env": s[45] / max(1, s[40]),
        },
        "raw": s,
    }
    txt = json.dumps(rec)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")
    env.close()


if __name__ == "__main__":
    main()
