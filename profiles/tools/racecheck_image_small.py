"""compute-sanitizer --tool racecheck target: 8 envs, reset + 24 image steps (trajectory boxes, span records, both image halves)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
n = 8
env = BatchedParkingEnv(n, scenes=generate_scenes(n, "mix", 5), auto_reset=True, use_img_observation=True)
env.reset()
rng = np.random.default_rng(0)
for _ in range(int(os.environ.get("STEPS", "24"))):
    obs, _, _, _ = env.step(torch.as_tensor(rng.uniform(-1, 1, size=(n, 2)), device=env.device).contiguous())
torch.cuda.synchronize()
print("image steps ok", int(obs["img"].sum()))
env.close()
