#!/usr/bin/env python
"""A/B of two builds of the kernels on the BASELINE cfg-3 workload, in one process on one GPU.

Steps the in-tree library (A) and a build with extra -D switches (B) over the same scenes with the same actions,
(1) compares EVERY output array of every step bit for bit, (2) times both with CUDA events in alternating blocks.
One JSON object per stage is printed as soon as it is known (a cut-off run still leaves the earlier ones).

Usage (GPU box):  python profiles/tools/ab_compare.py --defs=-DHOPE_CHK_PAIR=1 --name pair [--envs 65536]
                  python profiles/tools/ab_compare.py --variants "pair=-DHOPE_CHK_PAIR=1;persist=-DHOPE_OBS_PERSISTENT=32"
                  python profiles/tools/ab_compare.py --variants ... --build-only     (here, no GPU needed; the .so files travel)
"""
import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)


def build_variant(name, defs, force=False):
    from hope_b200 import build as hb
    defs = " ".join(tok for tok in defs.split() if not tok.startswith("ENV:"))
    path = os.path.join(HERE, "_ab", f"libhope_b200_{name}.so")
    deps = [os.path.join(hb.CSRC, d) for d in hb.DEPS]
    if not force and os.path.exists(path) and all(os.path.getmtime(d) <= os.path.getmtime(path) for d in deps):
        return path
    os.makedirs(os.path.dirname(path), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([hb.nvcc_path()] + hb.NVCC_FLAGS + defs.split() + ["-o", path] + [os.path.join(hb.CSRC, s) for s in hb.SOURCES], env=env)
    return path


def say(out, **kw):
    txt = json.dumps(kw)
    print(txt, flush=True)
    if out:
        with open(out, "a") as f:
            f.write(txt + "\n")


def run_variant(args, name, defs, env_a, scenes, act, total):
    import torch
    from hope_b200 import build as hb, capi
    from hope_b200.batched_env import BatchedParkingEnv
    n = args.envs
    env_sets = [tok[4:].split("=", 1) for tok in defs.split() if tok.startswith("ENV:")]   # "ENV:NAME=VALUE": set while the variant's context is created
    defs = " ".join(tok for tok in defs.split() if not tok.startswith("ENV:"))
    path_b = build_variant(name, defs)
    path_a = hb.VARIANTS[16]
    hb.VARIANTS[16] = path_b
    capi._LIBS.pop(16, None)
    saved = {k: os.environ.get(k) for k, _ in env_sets}
    try:
        for k, v in env_sets:
            os.environ[k] = v
        env_b = BatchedParkingEnv(n, scenes=scenes, auto_reset=True)  # binds the variant library
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        hb.VARIANTS[16] = path_a
        capi._LIBS[16] = env_a.lib
    assert env_a.lib is not env_b.lib
    say(args.out, stage="loaded", variant=name, a=os.path.basename(path_a), b=os.path.basename(path_b), defs=defs, envs=n)

    def bits(t):
        return t.view(torch.int64) if t.dtype == torch.float64 else t

    env_a.reset(); env_b.reset()
    ca0, cb0 = env_a.counters(), env_b.counters()
    mismatches, compared = {}, 0
    for k in range(total):
        env_a.step(act[k]); env_b.step(act[k])
        for field in env_a.out:
            if not torch.equal(bits(env_a.out[field]), bits(env_b.out[field])):
                mismatches[field] = mismatches.get(field, 0) + int((bits(env_a.out[field]) != bits(env_b.out[field])).sum().item())
        compared += 1
    ca, cb = env_a.counters(), env_b.counters()
    say(args.out, stage="parity", variant=name, steps_compared=compared, fields=len(env_a.out), mismatching_elements=mismatches,
        env_steps_a=ca["env_steps"] - ca0["env_steps"], env_steps_b=cb["env_steps"] - cb0["env_steps"],
        identical=(not mismatches and ca["env_steps"] - ca0["env_steps"] == cb["env_steps"] - cb0["env_steps"]))

    def timed(env, k0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for k in range(args.time_steps):
            env.step(act[(k0 + k) % total])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.time_steps

    ms_a, ms_b = [], []
    for blk in range(args.blocks):
        ms_a.append(timed(env_a, blk)); ms_b.append(timed(env_b, blk))
    say(args.out, stage="timing", variant=name, ms_per_step_a=ms_a, ms_per_step_b=ms_b, a_over_b=[x / y for x, y in zip(ms_a, ms_b)],
        note="device-resident hope_step, CUDA events, alternating blocks; not a bench value (two envs resident, no clock sampling)")
    for which, env in (("a", env_a), ("b", env_b)):  # per-kernel CUDA-event times inside the live step
        env.profile(True); env.profile_read()
        for k in range(20):
            env.step(act[k % total])
        pr = env.profile_read(); env.profile(False)
        say(args.out, stage="kernels_" + which, variant=name, ms_per_launch={kn: (v[0] / v[1] if v[1] else None) for kn, v in pr.items()})
    for which, env in (("a", env_a), ("b", env_b), ("a", env_a), ("b", env_b)):  # each kernel alone (no observe / Reeds-Shepp overlap): the figure to compare kernels on
        env.profile(True, serial=True); env.profile_read()
        for k in range(20):
            env.step(act[k % total])
        pr = env.profile_read(); env.profile(False)
        say(args.out, stage="kernels_serial_" + which, variant=name, ms_per_launch={kn: (v[0] / v[1] if v[1] else None) for kn, v in pr.items()})
    env_b.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--defs", default=None, help="one variant: its -D switches (with --name)")
    ap.add_argument("--name", default=None)
    ap.add_argument("--variants", default=None, help="several variants: 'name=-DX=1 -DY=2;other=-DZ=1'")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--check-steps", type=int, default=25)
    ap.add_argument("--time-steps", type=int, default=50)
    ap.add_argument("--blocks", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    variants = []
    if args.variants:
        variants += [tuple(v.split("=", 1)) for v in args.variants.split(";") if v.strip()]
    if args.defs:
        variants.append((args.name or "variant", args.defs))
    assert variants, "give --defs/--name or --variants"
    if args.build_only:
        for name, defs in variants:
            print(build_variant(name, defs))
        return
    import torch
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n = args.envs
    scenes = generate_scenes(2 * n, "mix", 42)
    env_a = BatchedParkingEnv(n, scenes=scenes, auto_reset=True)
    gen = torch.Generator(device=env_a.device); gen.manual_seed(1234)
    total = args.warmup + args.check_steps
    act = torch.rand((total, n, 2), dtype=torch.float64, device=env_a.device, generator=gen) * 2 - 1
    for name, defs in variants:
        run_variant(args, name, defs, env_a, scenes, act, total)
    env_a.close()


if __name__ == "__main__":
    main()
