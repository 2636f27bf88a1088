"""The drop-in boundary: what the reference's own scripts import from `env.*` resolves against the facade,
the backend reads the caller's `configs` module, and the host-side pieces that sit next to the step — Dragon Lake Parking
scene preparation and the `get_map_level` classifier — reproduce recordings of the unmodified reference.

CPU only (no CUDA call): the facade's modules import without a GPU; constructing an env does not.
"""
import ast
import glob
import importlib
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "hope_b200", "compat")
REF_SRC = "/root/reference/src"
LEVELS = ["Normal", "Complex", "Extrem"]


def _purge(prefixes=("env", "configs")):
    for m in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in prefixes)]:
        del sys.modules[m]
    from hope_b200 import refconfig
    refconfig._CACHE.clear()


@pytest.fixture()
def facade():
    """`env` resolves to hope_b200/compat/env, as with PYTHONPATH=<repo>/hope_b200/compat:<repo> (compat/env/__init__.py)"""
    _purge()
    sys.path.insert(0, COMPAT)
    yield
    sys.path.remove(COMPAT)
    _purge()


def _env_imports(path):
    """(module, [names]) of every `from env.X import ...` / `import env.X` statement of a script"""
    tree = ast.parse(open(path).read(), filename=path)
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and (node.module == "env" or node.module.startswith("env.")):
            found.append((node.module, [a.name for a in node.names]))
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name == "env" or a.name.startswith("env."):
                    found.append((a.name, []))
    return found


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_every_env_import_of_the_reference_scripts_resolves_against_the_facade(facade):
    """src/train/*.py and src/evaluation/*.py run from src/ and append '..' and '.' to sys.path (train_HOPE_sac.py:2-3), so a
    PYTHONPATH entry in front shadows their `env` package: every name they import from it must exist in the facade."""
    scripts = sorted(glob.glob(os.path.join(REF_SRC, "train", "*.py")) + glob.glob(os.path.join(REF_SRC, "evaluation", "*.py")))
    assert len(scripts) >= 4
    seen = set()
    for path in scripts:
        for module, names in _env_imports(path):
            mod = importlib.import_module(module)
            assert os.path.realpath(mod.__file__).startswith(os.path.realpath(COMPAT)), f"{module} resolved outside the facade: {mod.__file__}"
            for name in names:
                if name == "*":
                    continue
                assert hasattr(mod, name), f"{os.path.relpath(path, REF_SRC)}: `from {module} import {name}` does not resolve against the facade"
                seen.add((module, name))
    # the complete list as of the reference's commit 2accab9: a new import there shows up as a failure above, a stale list here
    assert seen == {("env.car_parking_base", "CarParking"), ("env.env_wrapper", "CarParkingWrapper"), ("env.vehicle", "VALID_SPEED"),
                    ("env.vehicle", "Status"), ("env.map_level", "get_map_level")}


def test_facade_modules_carry_the_reference_surface(facade):
    """attributes the callers touch, checked without the reference tree"""
    import env.car_parking_base as cpb
    import env.env_wrapper as wrap
    import env.map_base as map_base
    import env.map_level as map_level
    import env.vehicle as vehicle
    assert [s.name for s in vehicle.Status] == ["CONTINUE", "ARRIVED", "COLLIDED", "OUTBOUND", "OUTTIME"] and vehicle.Status.OUTTIME.value == 5
    assert vehicle.VALID_SPEED == [-2.5, 2.5] and vehicle.VALID_STEER == [-0.75, 0.75] and vehicle.NUM_STEP == 10
    v = vehicle.Vehicle()
    assert v.kinetic_model.step_len * v.kinetic_model.n_step * vehicle.VALID_SPEED[1] == 1.25  # train_HOPE_sac.py:164
    st = vehicle.State([1.0, 2.0, 0.5])
    assert st.get_pos() == (1.0, 2.0, 0.5) and st.loc.distance(vehicle.State([4.0, 6.0, 0.0]).loc) == 5.0
    box = st.create_box()
    assert len(box.coords) == 5 and box.coords[0] == box.coords[-1]
    for name in ("reset", "step", "render", "set_level", "close", "coord_transform_matrix"):
        assert callable(getattr(cpb.CarParking, name))
    for name in ("reward_shaping", "action_rescale", "observation_rescale", "CarParkingWrapper"):
        assert hasattr(wrap, name)
    assert map_level.get_map_level(st, vehicle.State([0, 0, 0]), []) == "Normal"
    assert map_base.Area(shape=box).get_shape().shape == (5, 2)


def test_wrapper_functions_equal_the_reference_expressions(facade):
    """the host versions of env_wrapper.py:10-55 the wrapper falls back to for custom hooks"""
    import env.env_wrapper as wrap
    from env.car_parking_base import Box
    from env.vehicle import Status
    space = Box(np.array([-0.75, -2.5]).astype(np.float32), np.array([0.75, 2.5]).astype(np.float32))
    a = wrap.action_rescale(np.array([0.3, -1.7]), space)
    assert a.dtype == np.float64 and np.array_equal(a, [0.3 * 0.75, -2.5])
    info = {}
    ri = dict(time_cost=-0.01, rs_dist_reward=9.0, dist_reward=0.02, angle_reward=7.0, box_union_reward=0.001)
    _, r, _, info = wrap.reward_shaping(None, ri, Status.CONTINUE, info)
    assert r == (0 + 1 * -0.01 + 0 * 9.0 + 5 * 0.02 + 0 * 7.0 + 10 * 0.001) * 0.1 and info["status"] == Status.CONTINUE
    assert [wrap.reward_shaping(None, ri, s, {})[1] for s in (Status.ARRIVED, Status.COLLIDED, Status.OUTBOUND, Status.OUTTIME)] == [5.0, -5.0, -5.0, -0.1]
    img = np.zeros((64, 64, 3))
    assert wrap.observation_rescale({"img": img})["img"].shape == (3, 64, 64)


# ---- configs.py keeps taking effect -----------------------------------------------------------------------------------------
def _fake_configs(**over):
    from hope_b200 import refconfig
    d = refconfig.defaults()
    mod = types.ModuleType("configs")
    for k in refconfig._NAMES:
        setattr(mod, k, getattr(d, k))
    for k, v in over.items():
        setattr(mod, k, v)
    mod.__file__ = "<test configs>"
    return mod


def test_backend_reads_the_callers_configs_module(facade):
    from hope_b200 import capi, refconfig, tables
    sys.modules["configs"] = _fake_configs(TOLERANT_TIME=50, WHEEL_BASE=2.9, LIDAR_RANGE=12.0, ENV_COLLIDE=True, VALID_SPEED=[-2.0, 2.0],
                                           REWARD_WEIGHT={"time_cost": 2, "rs_dist_reward": 0, "dist_reward": 4, "angle_reward": 1, "box_union_reward": 8})
    cfg = refconfig.load(refresh=True)
    assert cfg.source == "<test configs>" and cfg.TOLERANT_TIME == 50
    p = refconfig.step_params(cfg)
    assert p["tolerant_time"] == 50 and p["wheel_base"] == 2.9 and p["lidar_range"] == 12.0 and p["env_collide"] == 1
    assert p["valid_speed"] == [-2.0, 2.0] and p["reward_weight"] == [2.0, 0.0, 4.0, 1.0, 8.0]
    assert abs(p["box_x"][1] - (0.96 + 2.9)) < 1e-15  # VehicleBox follows WHEEL_BASE (configs.py:20-24)
    tb = tables.host_tables(cfg)
    base = tables.host_tables(refconfig.defaults())
    assert tb["dist_star"].shape == base["dist_star"].shape and not np.array_equal(tb["dist_star"], base["dist_star"])
    assert not np.array_equal(tb["lidar_base"], base["lidar_base"])
    import env.vehicle as vehicle
    import env.env_wrapper as wrap
    assert vehicle.VALID_SPEED == [-2.0, 2.0] and wrap.REWARD_WEIGHT["dist_reward"] == 4
    # what the compiled kernels cannot do is refused by name, not ignored
    sys.modules["configs"] = _fake_configs(LIDAR_NUM=90)
    with pytest.raises(capi.HopeError, match="LIDAR_NUM"):
        refconfig.validate(refconfig.load(refresh=True))
    sys.modules["configs"] = _fake_configs(PRECISION=5)
    with pytest.raises(capi.HopeError, match="discrete actions"):
        refconfig.validate(refconfig.load(refresh=True))
    del sys.modules["configs"]
    assert refconfig.load(refresh=True).source.startswith("reference defaults")


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_defaults_equal_the_reference_configs_module():
    """refconfig.defaults() against the unmodified src/configs.py (imported with the shapely stand-in)"""
    from hope_b200 import refconfig
    _purge()
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    try:
        ref = importlib.import_module("configs")
        assert ref.__file__.startswith(REF_SRC)
        snap, d = refconfig.load(refresh=True), refconfig.defaults()
        assert snap.source == ref.__file__
        for name in refconfig._NAMES:
            a, b = getattr(snap, name), getattr(d, name)
            assert (list(a.items()) == list(b.items())) if isinstance(a, dict) else (a == b), name
        assert np.array_equal(snap.VEHICLE_BOX, d.VEHICLE_BOX) and np.array_equal(np.array(snap.discrete_actions), np.array(ref.discrete_actions))
        assert np.array_equal(refconfig.palette(snap)[5:], np.array(ref.TRAJ_COLORS)[:, :3])
    finally:
        sys.path.remove(REF_SRC); sys.path.remove(os.path.join(ROOT, "oracle", "refshim"))
        _purge(("env", "configs", "shapely"))


# ---- Dragon Lake Parking preparation and get_map_level, pinned on the reference (oracle/make_dlp_golden.py) ---------------------
@pytest.fixture(scope="module")
def dlp_golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "dlp_reset.npz")))


class _GlobalRng(object):  # numpy's global generator, the one ParkingMapDLP.reset draws from
    integers = staticmethod(lambda lo, hi: np.random.randint(lo, hi))
    standard_normal = staticmethod(lambda n: np.array([np.random.randn() for _ in range(n)]))
    random = staticmethod(lambda: np.random.random())


def test_dlp_reset_reproduces_the_reference_from_the_same_seed(facade, golden_dir, dlp_golden):
    """ParkingMapDLP.reset (parking_map_dlp.py:38-86) recorded from the unmodified reference under np.random.seed(s): the case drawn,
    the jittered start, bounds, which obstacles survive the filter, the flips and the map level — bit for bit."""
    from hope_b200 import dlp
    from env.map_level import get_map_level
    g = dlp_golden
    cases = dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))
    code = {"Normal": 0, "Complex": 1, "Extrem": 2}
    flips = 0
    for r in range(len(g["seed"])):
        np.random.seed(int(g["seed"][r]))
        arg = int(g["case_arg"][r])
        cid = int(np.random.randint(0, len(cases))) if arg < 0 else (arg % len(cases) if arg >= len(cases) else arg)
        assert cid == g["case_id"][r]
        sc = dlp.prepare_scene(cases[cid], _GlobalRng)
        assert np.array_equal(sc["start"], g["start"][r]) and np.array_equal(sc["dest"], g["dest"][r]), r
        assert np.array_equal(sc["bounds"], g["bounds"][r])
        k = int((sc["nverts"] > 0).sum())
        assert k == g["n_obst"][r]
        kept = [j for j in range(len(cases[cid]["rings"])) if g["kept"][r][j]]
        assert all(np.array_equal(sc["obs"][i, :sc["nverts"][i]], cases[cid]["rings"][j]) for i, j in enumerate(kept))
        rings = [sc["obs"][i, :sc["nverts"][i]] for i in range(k)]
        assert code[get_map_level(sc["start"], sc["dest"], rings)] == g["level"][r], r
        flips += int(not np.array_equal(sc["dest"], cases[cid]["dest"]))
    assert 20 < flips < 100  # both branches of the dest flip were recorded


@pytest.mark.skipif(not os.path.exists("/root/reference/data/dlp.data"), reason="reference data file not present (GPU box)")
def test_get_map_level_labels_all_248_dlp_cases_like_the_reference(facade, dlp_golden):
    from hope_b200 import dlp
    from env.map_level import get_map_level
    cases = dlp.read_dlp("/root/reference/data/dlp.data")
    code = {"Normal": 0, "Complex": 1, "Extrem": 2}
    got = np.array([code[get_map_level(c["starts"][0], c["dest"], c["rings"])] for c in cases])
    assert np.array_equal(got, dlp_golden["levels_all"]), np.flatnonzero(got != dlp_golden["levels_all"])
    assert len(set(got)) == 3


def test_get_map_level_on_generated_scenes(facade):
    """generated bay / parallel scenes classify by construction: Extrem-level parallel lots are shorter than
    EXTREM_PARK_LOT_LENGTH (map_level.py:11), Normal-level scenes never label Extrem"""
    from hope_b200.batched_env import generate_scenes
    from env.map_level import get_map_level
    labels = {}
    for lv in LEVELS:
        sc = generate_scenes(40, lv, 7)
        labels[lv] = [get_map_level(sc["start"][i], sc["dest"][i], [sc["obs"][i, k, :nv] for k, nv in enumerate(sc["nverts"][i]) if nv])
                      for i in range(40)]
    assert "Extrem" not in labels["Normal"] and labels["Extrem"].count("Extrem") >= 30
    assert labels["Normal"].count("Normal") > labels["Complex"].count("Normal")


def test_dlp_reader_refuses_foreign_pickles(tmp_path):
    """the unpickler of hope_b200.dlp allows the three globals of data/dlp.data and nothing else"""
    import pickle
    from hope_b200 import dlp
    p = tmp_path / "evil.data"
    p.write_bytes(pickle.dumps([(eval, ("1+1",))]))
    with pytest.raises(pickle.UnpicklingError, match="refusing"):
        dlp.read_dlp(str(p))


def test_trajectory_rendering_settings_follow_configs():
    """configs.py:86 TRAJ_RENDER_LEN and :105 RENDER_TRAJ reach k_render as a run-time length (hope_set_render_traj); the palette
    keeps its 25 rows, the trajectory colours of a shorter trail in rows 5 .. 5 + len - 1 (configs.py:87-88 TRAJ_COLORS)."""
    from hope_b200 import refconfig
    c = refconfig.defaults()
    assert refconfig.traj_render_len(c) == 20 and refconfig.image_supported(c)
    pal20 = refconfig.palette(c)
    assert pal20.shape == (25, 3) and tuple(pal20[5]) == (10, 10, 10) and tuple(pal20[24]) == (10, 10, 200)
    c.TRAJ_RENDER_LEN = 7
    pal7 = refconfig.palette(c)
    want = np.linspace(np.array(c.TRAJ_COLOR_LOW), np.array(c.TRAJ_COLOR_HIGH), 7, endpoint=True, dtype=np.uint8)[:, :3]
    assert refconfig.traj_render_len(c) == 7 and refconfig.image_supported(c)
    assert pal7.shape == (25, 3) and np.array_equal(pal7[5:12], want) and np.array_equal(pal7[:5], pal20[:5])
    c.RENDER_TRAJ = False
    assert refconfig.traj_render_len(c) == 0 and refconfig.image_supported(c)
    c.RENDER_TRAJ, c.TRAJ_RENDER_LEN = True, 21
    assert not refconfig.image_supported(c)   # the trajectory ring buffer of k_advance holds 20 poses
    c.TRAJ_RENDER_LEN, c.K = 20, 10
    assert not refconfig.image_supported(c)   # the raster geometry itself stays compile-time
