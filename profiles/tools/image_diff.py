"""Debug aid: step a few generated scenes with the image stage on, compare k_render with oracle/image_oracle.py and
dump the differing images into gpurun_out/image_diff.npz (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes  # noqa: E402
from oracle import image_oracle as io  # noqa: E402

n, steps = 24, 30
sc = generate_scenes(n, "mix", 31)
env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
book = io.TrajectoryBook(n)
obs = env.reset()
for i in range(n):
    book.reset(i, sc["start"][i])
rng = np.random.default_rng(5)
drift = rng.uniform(-1, 1, size=(n, 2))
bad = []
tot = 0
for k in range(-1, steps):
    if k >= 0:
        act = np.clip(0.7 * drift + 0.3 * rng.uniform(-1, 1, size=(n, 2)), -1, 1)
        obs, _, done, _ = env.step(torch.as_tensor(act, device=env.device).contiguous())
        pose, sub, ret = (env.out[x].cpu().numpy() for x in ("pose", "substeps", "retreated"))
        for i in range(n):
            book.step(i, pose[i], sub[i], ret[i])
    got = obs["img"].cpu().numpy()
    for i in range(0, n, 2):
        rings = io.scene_rings(sc["obs"][i], sc["nverts"][i])
        want = io.render_observation(sc["start"][i], sc["dest"][i], sc["bounds"][i], rings, book.traj[i])
        tot += 1
        if not np.array_equal(got[i], want):
            bad.append((k, i, got[i].copy(), want, int((got[i] != want).sum())))
print("compared", tot, "differing", len(bad), [(b[0], b[1], b[4]) for b in bad[:20]])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if bad:
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "image_diff.npz"), step=[b[0] for b in bad[:16]], env=[b[1] for b in bad[:16]],
                        got=np.stack([b[2] for b in bad[:16]]), want=np.stack([b[3] for b in bad[:16]]),
                        traj=np.array([book.traj[b[1]][-1] for b in bad[:16]]))
