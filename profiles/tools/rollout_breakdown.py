#!/usr/bin/env python
"""Where the cfg-4 rollout step goes: CUDA-event time of each stage of RolloutEngine.act + the env step (65 536 envs)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
from hope_b200 import rollout
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes


def timed(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--image", action="store_true")
    args = ap.parse_args()
    n = args.envs
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", 1), auto_reset=True, use_img_observation=args.image)
    dev = env.device
    actor, what = rollout.reference_actor(use_img=args.image, device=dev)
    out = {"policy": what, "envs": n}
    for label, kw in (("fused+graph", dict(fused=True, graph=True)), ("fused", dict(fused=True, graph=False)), ("eager", dict(fused=False))):
        eng = rollout.RolloutEngine(env, actor, seed=0, **kw)
        eng.collect(4)
        obs = eng.obs
        r = {"step_total": timed(lambda: eng.collect(1))}
        r["env_step"] = timed(lambda: env.step(eng.sampler.action if eng.fused else torch.zeros((n, 2), dtype=torch.float64, device=dev)))
        r["act"] = timed(lambda: eng.act(obs))
        if eng.fused:
            r["norm_kernel"] = timed(lambda: eng.norm(obs))
            net_in = dict(eng.norm.out)
            if args.image:
                net_in["img"] = eng._img_f32
            r["policy_forward"] = timed(lambda: eng._policy_mean(net_in))
            m32 = eng._policy_mean(net_in).contiguous()
            r["sample_kernel"] = timed(lambda: eng.sampler(m32, eng.log_std, obs["action_mask"]))
        r["planner"] = timed(lambda: env.planner_actions(torch.zeros((n, 2), dtype=torch.float64, device=dev)))
        out[label] = r
    print(json.dumps(out))
    env.close()


if __name__ == "__main__":
    main()
