"""compute-sanitizer target: smoke() (device path, image on) plus two host-API steps (narrow mask format, host threads)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
g.smoke()
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
env = BatchedParkingEnv(9000, scenes=generate_scenes(9000, "mix", 5), auto_reset=True, use_img_observation=True)
env.reset_host()
rng = np.random.default_rng(0)
for _ in range(2):
    h = env.step_host(rng.uniform(-1, 1, size=(9000, 2)), outputs=BatchedParkingEnv.HOST_DEFAULT + ("img",))
print("host steps ok", float(h["mask"].sum()), int(h["img"].sum()))
env.close()
