#!/usr/bin/env python
"""hope_step_host with one process per GPU of a box (torchrun), several wire-format settings in one launch.

Every rank steps its own 65 536 envs through the host API; a setting's time is the max over ranks of the wall clock around K
synchronous steps between two barriers (gloo).  The settings are environment variables read at hope_create, so each one gets a
fresh context in the same process.  Not a bench value (no clock sampling): it ranks the settings at this rank count.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      profiles/tools/e2e_ranks.py --out gpurun_out/e2e_ranks.jsonl "" "HOPE_B200_HOST_PACK_FRAC=0.25" "HOPE_B200_HOST_LIDAR_PACK=0"
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("settings", nargs="*")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    scenes = generate_scenes(2 * args.envs, "mix", 42 + 7919 * rank)
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    act = (torch.rand((args.steps + args.warmup, args.envs, 2), dtype=torch.float64, device=dev, generator=gen) * 2 - 1).cpu().numpy()
    base = dict(os.environ)
    for setting in (args.settings or [""]):
        os.environ.clear(); os.environ.update(base)
        for kv in setting.split():
            k, v = kv.split("=", 1)
            os.environ[k] = v
        env = BatchedParkingEnv(args.envs, scenes=scenes, device=local, auto_reset=True)
        env.reset()
        for k in range(args.warmup):
            env.step_host(act[k])
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(args.warmup, args.warmup + args.steps):
            env.step_host(act[k])
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64)
        tmin = t.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        w = env.host_wire_info()
        env.close()
        if rank == 0:
            line = json.dumps({"env": setting, "ranks": world, "ms_per_step_max_over_ranks": 1e3 * float(t) / args.steps, "ms_per_step_fastest_rank": 1e3 * float(tmin) / args.steps,
                               "env_steps_per_s_nominal": world * args.envs * args.steps / float(t), "wire_rank0": w})
            print(line, flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(line + "\n")
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
