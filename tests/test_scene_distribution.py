"""The product's scene generators against the reference's, as distributions (the streams differ, so scenes cannot be
compared one by one): two-sample Kolmogorov-Smirnov distance of 15 scalar features per level and case type between 3 000
product scenes and 1 500 scenes of the unmodified `ParkingMapNormal.reset` (tests/golden/scene_stats.npz, recorded by
oracle/make_scene_stats.py with the very `features()` function used here).  parking_map_normal.py:40-494."""
import os

import numpy as np
import pytest

from hope_b200.batched_env import generate_scenes
from oracle.make_scene_stats import FEATURES, features

LEVELS = ("Normal", "Complex", "Extrem")
N = 3000


def ks_distance(a, b):
    a, b = np.sort(a), np.sort(b)
    grid = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, grid, side="right") / len(a) - np.searchsorted(b, grid, side="right") / len(b)).max())


def critical(n, m, c_alpha=2.2):
    """rejection threshold of the two-sample KS test; c = 2.2 is alpha ~ 1e-4, so that ~90 comparisons pass together"""
    return c_alpha * np.sqrt((n + m) / (n * m))


def compare(gold, level, sc):
    rows = {0: [], 1: []}
    for i in range(len(sc["start"])):
        rows[int(sc["case_id"][i])].append(features(sc["start"][i], sc["dest"][i], sc["bounds"][i], sc["obs"][i], sc["nverts"][i]))
    bay = len(rows[0]) / len(sc["start"])
    want_bay = float(gold[f"{level}_bay_fraction"])
    assert abs(bay - want_bay) < 0.04, (level, bay, want_bay)  # parking_map_normal.py:475-480: p = 0.5 for Normal / Complex, 0 for Extrem
    worst = []
    for case, name in ((0, "bay"), (1, "parallel")):
        if not rows[case]:
            continue
        for f in FEATURES:
            ref = gold[f"{level}_{name}_{f}"]
            got = np.array([r[f] for r in rows[case]])
            d, lim = ks_distance(got, ref), critical(len(got), len(ref))
            worst.append((d / lim, f"{level}/{name}/{f}", d, lim, float(np.mean(got)), float(np.mean(ref))))
    worst.sort(reverse=True)
    return worst


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "scene_stats.npz")))


@pytest.mark.parametrize("level", LEVELS)
def test_host_generator_matches_the_reference_distribution(gold, level):
    sc = generate_scenes(N, level, 31337)
    worst = compare(gold, level, sc)
    print("\n" + "\n".join(f"  {w[1]:42s} KS {w[2]:.3f} (limit {w[3]:.3f})  mean {w[4]:8.3f} vs reference {w[5]:8.3f}" for w in worst[:6]))
    assert worst[0][0] < 1.0, worst[0]


@pytest.mark.gpu
@pytest.mark.parametrize("level", LEVELS)
def test_device_generator_matches_the_reference_distribution(gold, level):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hope_b200.batched_env import BatchedParkingEnv
    env = BatchedParkingEnv(N, pool_size=N, level=level, seed=4711, auto_reset=False, device_scenes=True)
    sc = env.get_scene_pool()
    # the device pool does not report the case type: bay lots sit on the y = 0 wall with the slot pointing up
    # (parking_map_normal.py:62-66, 74-80: dest yaw ~ pi/2), parallel lots have dest yaw ~ 0 or pi
    sc["case_id"] = (np.abs(np.sin(sc["dest"][:, 2])) < 0.7).astype(np.int32)
    worst = compare(gold, level, sc)
    print("\n" + "\n".join(f"  {w[1]:42s} KS {w[2]:.3f} (limit {w[3]:.3f})  mean {w[4]:8.3f} vs reference {w[5]:8.3f}" for w in worst[:6]))
    assert worst[0][0] < 1.0, worst[0]
    env.close()
