"""GPU parity of the image observation: k_render through the C ABI against
(1) the uint8 images recorded from the unmodified reference (tests/golden/images_*.npz) and
(2) oracle/image_oracle.py on generated and Dragon-Lake scenes.  Bar: bit-exact bytes (integer work).

What is pinned: the reference's own rendering glue and the real cv2.resize; the pygame raster rules are a
restatement (oracle/softraster.py), so parity with a real pygame build is unpinned."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from hope_b200 import capi  # noqa: E402
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes  # noqa: E402
from oracle import image_oracle as io  # noqa: E402

LEVELS = ("Normal", "Complex", "Extrem")


def _np(t):
    return t.detach().cpu().numpy()


def _mismatch(a, b):
    return int((a != b).any(axis=(1, 2, 3)).sum()), int((a != b).sum())


@pytest.mark.parametrize("level", LEVELS)
def test_golden_images_through_cuda(golden_dir, level):
    g = dict(np.load(os.path.join(golden_dir, f"images_{level}.npz")))
    n_ep = len(g["scene_start"])
    scenes = dict(start=g["scene_start"], dest=g["scene_dest"], bounds=g["scene_bounds"], obs=g["scene_obs"], nverts=g["scene_nverts"])
    env = BatchedParkingEnv(n_ep, scenes=scenes, auto_reset=False, use_img_observation=True)
    obs = env.reset()
    assert obs["img"].shape == (n_ep, 3, 64, 64) and obs["img"].dtype == torch.uint8
    assert np.array_equal(_np(obs["img"]), g["scene_reset_img"]), "reset images"
    idx = [np.where(g["ep"] == e)[0] for e in range(n_ep)]
    checked = 0
    for k in range(max(len(i) for i in idx)):
        live = np.array([k < len(i) for i in idx])
        rows = np.array([i[k] if k < len(i) else i[-1] for i in idx])
        act = np.where(live[:, None], g["action"][rows], 0.0)
        obs, _, _, _ = env.step(torch.as_tensor(act, device=env.device).contiguous())
        img = _np(obs["img"])
        assert np.array_equal(_np(env.out["substeps"])[live], g["substeps"][rows][live])
        assert np.array_equal(img[live], g["img"][rows][live]), f"step {k}: {_mismatch(img[live], g['img'][rows][live])}"
        checked += int(live.sum())
    assert checked == len(g["ep"])
    env.close()


def _oracle_images(sc, book, ids, traj_render_len=io.TRAJ_RENDER_LEN):
    out = []
    for i in ids:
        rings = io.scene_rings(sc["obs"][i], sc["nverts"][i])
        out.append(io.render_observation(sc["start"][i], sc["dest"][i], sc["bounds"][i], rings, book.traj[i], traj_render_len))
    return np.stack(out)


def _lockstep_images(env, sc, n, steps, seed, check_every=3, redraw=12, keep=0.7):
    """Steps the CUDA env with persistent, drifting actions; the oracle renders from the CUDA env's own pose and
    substep counters (so this isolates the rasteriser), a rotating subset of envs each step."""
    book = io.TrajectoryBook(n)
    obs = env.reset()
    for i in range(n):
        book.reset(i, sc["start"][i])
    ids = list(range(0, n, 4))
    assert np.array_equal(_np(obs["img"])[ids], _oracle_images(sc, book, ids)), "reset images"
    rng = np.random.default_rng(seed)
    drift = rng.uniform(-1, 1, size=(n, 2))
    live = np.ones(n, dtype=bool)
    compared = 0
    for k in range(steps):
        if k % redraw == redraw - 1:
            drift = rng.uniform(-1, 1, size=(n, 2))
        act = np.clip(keep * drift + (1.0 - keep) * rng.uniform(-1, 1, size=(n, 2)), -1, 1)
        obs, _, done, _ = env.step(torch.as_tensor(act, device=env.device).contiguous())
        pose, sub, ret = _np(env.out["pose"]), _np(env.out["substeps"]), _np(env.out["retreated"])
        for i in range(n):
            if live[i]:
                book.step(i, pose[i], sub[i], ret[i])
        ids = [i for i in range(k % check_every, n, check_every) if live[i]]
        if ids:
            got, want = _np(obs["img"])[ids], _oracle_images(sc, book, ids)
            assert np.array_equal(got, want), f"step {k}: (images, bytes) differing = {_mismatch(got, want)}"
            compared += len(ids)
        live &= _np(done) == 0
    return compared, book


def test_render_matches_the_oracle_on_generated_scenes():
    n = 48
    sc = generate_scenes(n, "mix", 31)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    compared, book = _lockstep_images(env, sc, n, 45, 5)
    assert compared >= 300
    assert max(len(t) for t in book.traj) > io.TRAJ_RENDER_LEN
    env.close()


def test_render_dynamic_layer_resolved_per_lattice_sample(monkeypatch):
    """k_render keeps the vehicle and trajectory boxes in a screen window of shared memory; a trail too long for the
    window is resolved per lattice sample instead.  HOPE_B200_RENDER_LATTICE=1 (read by hope_create) takes that path
    for every image."""
    monkeypatch.setenv("HOPE_B200_RENDER_LATTICE", "1")
    n = 32
    sc = generate_scenes(n, "mix", 33)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    monkeypatch.delenv("HOPE_B200_RENDER_LATTICE")
    compared, book = _lockstep_images(env, sc, n, 36, 6)
    assert compared >= 150
    assert max(len(t) for t in book.traj) > io.TRAJ_RENDER_LEN
    env.close()


def test_render_long_trails_at_full_speed():
    """Full speed held for the whole run (forward for one half of the envs, backward for the other, a little steering):
    trails of 20 boxes spanning more than 200 pixels on every heading, i.e. dynamic-layer windows near and over the
    shared-memory capacity (the per-lattice-sample path is taken where the window does not fit)."""
    n = 64
    sc = generate_scenes(n, "mix", 35)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    env.reset()
    book = io.TrajectoryBook(n)
    for i in range(n):
        book.reset(i, sc["start"][i])
    rng = np.random.default_rng(4)
    act = np.stack([rng.uniform(-0.35, 0.35, size=n), np.where(np.arange(n) % 2 == 0, 1.0, -1.0)], axis=1)  # (steer, speed)
    live = np.ones(n, dtype=bool)
    compared, longest = 0, 0.0
    for k in range(30):
        obs, _, done, _ = env.step(torch.as_tensor(act, device=env.device).contiguous())
        pose_k, sub, ret = _np(env.out["pose"]), _np(env.out["substeps"]), _np(env.out["retreated"])
        for i in range(n):
            if live[i]:
                book.step(i, pose_k[i], sub[i], ret[i])
                tail = np.asarray([p[:2] for p in book.traj[i][-io.TRAJ_RENDER_LEN:]])
                longest = max(longest, float(np.abs(tail.max(axis=0) - tail.min(axis=0)).max()))
        ids = [i for i in range(k % 2, n, 2) if live[i]]
        if ids:
            got, want = _np(obs["img"])[ids], _oracle_images(sc, book, ids)
            assert np.array_equal(got, want), f"step {k}: (images, bytes) differing = {_mismatch(got, want)}"
            compared += len(ids)
        live &= _np(done) == 0
    assert compared >= 300
    assert longest * 12 > 180, longest  # K = 12 pixels per metre
    env.close()


@pytest.mark.parametrize("length, render_traj", [(7, True), (20, False), (1, True)])
def test_trajectory_length_follows_configs(length, render_traj):
    """configs.py:86 TRAJ_RENDER_LEN / :105 RENDER_TRAJ are run-time settings of k_render (hope_set_render_traj): a shorter trail
    takes TRAJ_COLORS of that length, RENDER_TRAJ = False draws no trail at all (car_parking_base.py:315-320)."""
    from hope_b200 import refconfig
    cfg = refconfig.defaults()
    cfg.TRAJ_RENDER_LEN, cfg.RENDER_TRAJ = length, render_traj
    drawn = length if render_traj else 0
    n = 24
    sc = generate_scenes(n, "mix", 36)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True, config=cfg)
    book = io.TrajectoryBook(n)
    obs = env.reset()
    for i in range(n):
        book.reset(i, sc["start"][i])
    rng = np.random.default_rng(9)
    drift = rng.uniform(-1, 1, size=(n, 2))
    live = np.ones(n, dtype=bool)
    compared = 0
    for k in range(14):
        act = np.clip(0.8 * drift + 0.2 * rng.uniform(-1, 1, size=(n, 2)), -1, 1)
        obs, _, done, _ = env.step(torch.as_tensor(act, device=env.device).contiguous())
        pose, sub, ret = _np(env.out["pose"]), _np(env.out["substeps"]), _np(env.out["retreated"])
        for i in range(n):
            if live[i]:
                book.step(i, pose[i], sub[i], ret[i])
        ids = [i for i in range(k % 2, n, 2) if live[i]]
        if ids:
            got, want = _np(obs["img"])[ids], _oracle_images(sc, book, ids, drawn)
            assert np.array_equal(got, want), f"step {k}: (images, bytes) differing = {_mismatch(got, want)}"
            compared += len(ids)
        live &= _np(done) == 0
    assert compared >= 80
    env.close()


@pytest.mark.parametrize("regenerate", [False, True])
def test_static_screen_follows_auto_reset_and_regeneration(regenerate):
    """The static part of the screen is cached per env and keyed by (pool slot, regeneration count of the slot): an env that
    auto-resets onto the next pool scene, or whose slot is regenerated on the device, must get a fresh screen (and its
    trajectory records must not survive the reset)."""
    n = 96
    if regenerate:
        env = BatchedParkingEnv(n, pool_size=n, level="Normal", seed=5, auto_reset=True, device_scenes=True, use_img_observation=True)
    else:
        pool = generate_scenes(2 * n, "Normal", 11)
        env = BatchedParkingEnv(n, scenes=pool, auto_reset=True, use_img_observation=True)
    env.reset()
    pool = env.get_scene_pool()
    book = io.TrajectoryBook(n)
    sid = env.get_state()["scene_id"]
    for i in range(n):
        book.reset(i, pool["start"][sid[i]])
    act = torch.zeros((n, 2), dtype=torch.float64, device=env.device); act[:, 1] = 1.0  # straight ahead until something ends the episode
    act[:, 0] = torch.linspace(-0.5, 0.5, n, dtype=torch.float64, device=env.device)
    resets, compared = 0, 0
    for k in range(40):
        obs, _, _, _ = env.step(act)
        pose, sub, ret, was = (_np(env.out[key]) for key in ("pose", "substeps", "retreated", "was_reset"))
        if was.any():
            pool, sid = env.get_scene_pool(), env.get_state()["scene_id"]
        for i in range(n):
            if was[i]:
                book.reset(i, pose[i]); resets += 1
            else:
                book.step(i, pose[i], sub[i], ret[i])
        ids = [i for i in range(n) if was[i]][:6] + list(range(k % 8, n, 8))
        scenes = {key: pool[key][sid] for key in ("start", "dest", "bounds", "obs", "nverts")}
        got, want = _np(obs["img"])[ids], _oracle_images(scenes, book, ids)
        assert np.array_equal(got, want), f"step {k}: (images, bytes) differing = {_mismatch(got, want)}"
        compared += len(ids)
    assert resets >= n // 2 and compared >= 400
    env.close()


def test_quarter_turn_headings_take_the_rotate90_path():
    n = 8
    sc = generate_scenes(n, "Normal", 77)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    env.reset()
    book = io.TrajectoryBook(n)
    pose = sc["start"].copy()
    heads = [0.0, np.pi / 2, np.pi, -np.pi / 2, 3 * np.pi / 2, 2 * np.pi, 1e-9, -np.pi]
    for i in range(n):
        book.reset(i, sc["start"][i])
        pose[i, 2] = heads[i]
    env.set_state(pose=pose)
    zero = torch.zeros((n, 2), dtype=torch.float64, device=env.device)  # speed 0: the pose stays where it was put
    obs, _, _, _ = env.step(zero)
    got_pose = _np(env.out["pose"])
    assert np.array_equal(got_pose[:, 2], pose[:, 2])
    for i in range(n):
        book.step(i, got_pose[i], _np(env.out["substeps"])[i], _np(env.out["retreated"])[i])
    assert np.array_equal(_np(obs["img"]), _oracle_images(sc, book, range(n)))
    env.close()


def test_image_stage_is_optional_and_does_not_disturb_the_other_outputs():
    n = 64
    sc = generate_scenes(n, "mix", 9)
    a = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    b = BatchedParkingEnv(n, scenes=sc, auto_reset=False)
    assert b.reset()["img"] is None
    a.reset()
    rng = np.random.default_rng(3)
    for _ in range(6):
        act = torch.as_tensor(rng.uniform(-1, 1, size=(n, 2)), device=a.device).contiguous()
        a.step(act); b.step(act)
    for key in ("pose", "lidar", "mask_steps", "status", "rs_found", "reward"):
        assert torch.equal(a.out[key], b.out[key]), key
    host = a.step_host(rng.uniform(-1, 1, size=(n, 2)), outputs=("img", "status"))
    assert host["img"].shape == (n, 3, 64, 64) and host["img"].any()
    a.close(); b.close()


def test_render_on_dragon_lake_scenes_128_ring_build(golden_dir):
    from hope_b200 import dlp
    cases = dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))
    n = 12
    sc = dlp.prepare_scenes(cases, np.arange(n) % 16, seed=5)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    assert env.max_obs == 128
    compared, _ = _lockstep_images(env, sc, n, 8, 2, check_every=2)
    assert compared >= 30
    env.close()


def test_full_size_65536_images_placement_invariance_and_oracle_subset():
    """BASELINE cfg-3 size with the image on: 256 distinct scenes tiled over 65 536 envs with identical actions per
    tile -> every copy must produce the identical image (size-independent property), and a few envs are checked
    against the oracle."""
    n, period = 65536, 256
    base = generate_scenes(period, "mix", 404)
    sc = {k: np.concatenate([v] * (n // period)) for k, v in base.items()}
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    book = io.TrajectoryBook(period)
    obs = env.reset()
    for i in range(period):
        book.reset(i, base["start"][i])
    rng = np.random.default_rng(8)
    drift = rng.uniform(-1, 1, size=(period, 2))
    for k in range(6):
        act = np.clip(0.7 * drift + 0.3 * rng.uniform(-1, 1, size=(period, 2)), -1, 1)
        obs, _, _, _ = env.step(torch.as_tensor(np.tile(act, (n // period, 1)), device=env.device).contiguous())
        pose, sub, ret = _np(env.out["pose"][:period]), _np(env.out["substeps"][:period]), _np(env.out["retreated"][:period])
        for i in range(period):
            book.step(i, pose[i], sub[i], ret[i])
    img = obs["img"].view(n // period, period, 3, 64, 64)
    assert bool((img == img[0:1]).all()), "copies of one scene rendered differently"
    assert int(img[0].any(dim=(1, 2, 3)).sum()) == period  # nothing is blank
    ids = list(range(0, period, 16))
    assert np.array_equal(_np(img[-1])[ids], _oracle_images(base, book, ids))
    env.close()


def test_palette_override_and_capacity_guard():
    """hope_set_palette replaces configs.py's colours (painter's order); a vehicle box too large for k_render's
    per-box span table is refused instead of rendered wrongly."""
    n = 16
    sc = generate_scenes(n, "Normal", 12)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True)
    pal = np.full((capi.N_COLOR, 3), 50, dtype=np.uint8)
    pal[0] = (255, 255, 255)  # background stays white -> black
    env.set_palette(pal)
    img = _np(env.reset()["img"])
    assert img.max() == 50 and set(np.unique(img)) <= {0, 13, 25, 38, 50}   # (k * 50 + 2) >> 2 for k covered samples of 4
    env.close()
    big = BatchedParkingEnv(n, scenes=sc, auto_reset=False, use_img_observation=True,
                            params={"box_x": (-3.0, 9.0, 9.0, -3.0)})
    with pytest.raises(capi.HopeError, match="(?i)capacity|span table|exceeds"):
        big.reset()
    big.close()
