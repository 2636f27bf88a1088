"""Dragon Lake Parking scenes (scope row f3): reader, scene preparation, and step parity on the 128-ring builds."""
import os

import numpy as np
import pytest

from hope_b200 import dlp
from oracle import parking_oracle as po

REF_DATA = "/root/reference/data/dlp.data"


@pytest.fixture(scope="module")
def cases(golden_dir):
    return dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))


def test_fixture_matches_the_reference_file(golden_dir):
    if not os.path.exists(REF_DATA):
        pytest.skip("reference data file not present (GPU box)")
    full = dlp.read_dlp(REF_DATA)
    assert len(full) == 248
    npz = np.load(os.path.join(golden_dir, "dlp_cases.npz"))
    fx = dlp.cases_from_fixture(npz)
    for j, c in enumerate(npz["case_ids"]):
        assert np.array_equal(fx[j]["dest"], full[c]["dest"])
        assert np.array_equal(fx[j]["starts"], full[c]["starts"][:8])
        assert len(fx[j]["rings"]) == len(full[c]["rings"]) and all(np.array_equal(a, b) for a, b in zip(fx[j]["rings"], full[c]["rings"]))
    n_obst = [len(c["rings"]) for c in full]
    assert min(n_obst) == 162 and max(n_obst) == 312


def test_scene_preparation_follows_parking_map_dlp(cases):
    rng = np.random.default_rng(0)
    for case in cases:
        sc = dlp.prepare_scene(case, rng, start_index=0, flips=(False, False))
        lo = np.minimum(sc["start"][:2], sc["dest"][:2]); hi = np.maximum(sc["start"][:2], sc["dest"][:2])
        assert np.array_equal(sc["bounds"], [np.floor(lo[0] - 20), np.ceil(hi[0] + 20), np.floor(lo[1] - 20), np.ceil(hi[1] + 20)])
        k = int((sc["nverts"] > 0).sum())
        assert 20 <= k <= 128 and (sc["nverts"][:k] >= 3).all() and (sc["nverts"][k:] == 0).all()
        xmin, xmax, ymin, ymax = sc["bounds"]
        for r in range(k):
            v = sc["obs"][r, :sc["nverts"][r]]
            assert not (v[:, 0].max() <= xmin or v[:, 0].min() >= xmax or v[:, 1].max() <= ymin or v[:, 1].min() >= ymax)
        f = dlp.prepare_scene(case, np.random.default_rng(0), start_index=0, flips=(True, False))
        # flipping keeps the box where it is: the rear axle moves to the other end, heading turns by pi
        assert abs(f["dest"][2] - sc["dest"][2] - np.pi) < 1e-12
        assert abs(np.hypot(*(f["dest"][:2] - sc["dest"][:2])) - 2 * 1.415) < 1e-9


def test_oracle_128_ring_build_steps_dlp_scenes(cases):
    sc = dlp.prepare_scenes(cases, range(16), seed=3)
    env = po.OracleEnv(sc["start"], sc["dest"], sc["bounds"], sc["obs"], sc["nverts"], nthreads=2)
    out = env.reset_step()
    assert (out["status"] == 1).all()          # recorded start poses touch nothing
    assert (out["lidar"] < 10 - 1e-6).any()    # and the lot is full of parked cars
    rng = np.random.default_rng(0)
    for _ in range(5):
        out = env.step(rng.uniform(-1, 1, size=(16, 2)))
    assert np.isfinite(out["lidar"]).all() and out["rs_err"].sum() == 0


@pytest.mark.gpu
def test_cuda_128_ring_build_matches_the_oracle_on_dlp_scenes(cases):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hope_b200.batched_env import BatchedParkingEnv
    from tests.test_gpu_parity import Tally, compare_step, gather, assert_bars, check_rs_found, FLOAT_TOL
    ids = np.arange(96) % 16
    sc = dlp.prepare_scenes(cases, ids, seed=11)
    env = BatchedParkingEnv(96, scenes=sc, auto_reset=False)
    assert env.max_obs == 128
    orc = po.OracleEnv(sc["start"], sc["dest"], sc["bounds"], sc["obs"], sc["nverts"])
    env.reset(); ref = orc.reset_step(stages=1)
    out = gather(env)
    assert np.abs(out["lidar"] - ref["lidar"]).max() <= FLOAT_TOL
    assert np.array_equal(out["mask_steps"], ref["mask_steps"].astype(np.uint8))
    rng = np.random.default_rng(2)
    tl = Tally(); live = np.ones(96, dtype=bool)
    flips = []
    for k in range(60):
        act = rng.uniform(-1, 1, size=(96, 2))
        env.step(torch.as_tensor(act, device=env.device).contiguous())
        ref = orc.step(act)
        nf = compare_step(tl, gather(env), {**ref, "mask_steps": ref["mask_steps"].astype(np.uint8)}, orc.pose, live)
        flips += [(k, e) for e in np.flatnonzero(nf)]
        live &= ref["status"] == 1
    tl.report("DLP scenes (128-ring build), 96 envs x 60 steps")
    assert_bars(tl)
    check_rs_found("dlp:96x60", flips)
    env.close()
