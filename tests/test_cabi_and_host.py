"""CPU tests of the product's host side: the C ABI library loads and exports every symbol the
header declares (no compute calls without a GPU), struct layouts match, the host scene generator
and the constant tables behave, and the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "hope_b200.h")).read()


def test_library_exports_every_declared_symbol():
    from hope_b200 import capi
    lib = capi.load_library()
    declared = set(re.findall(r"\b(hope_[a-z_0-9]+)\s*\(", _header()))
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hope_b200.h but not exported"
    assert lib.hope_version() >= 100
    assert lib.hope_strerror(0) == b"ok" and b"invalid" in lib.hope_strerror(-1)


def test_ctypes_structs_match_the_header():
    from hope_b200 import capi
    h = _header()
    out_body = h[h.index("typedef struct hope_out {"):h.index("} hope_out;")]
    members = re.findall(r"\*(\w+);", out_body)
    assert members == [name for name, _, _ in capi.OUT_FIELDS]
    assert C.sizeof(capi.Out) == 8 * len(members)
    p = capi.Params()
    capi.check(capi.load_library().hope_default_params(C.byref(p)))
    # configs.py:13-38, 95-104, 180-187
    assert (p.wheel_base, p.num_step, p.step_length, p.mini_iter) == (2.8, 10, 0.05, 20)
    assert list(p.box_x) == [-0.93, 0.96 + 2.8, 0.96 + 2.8, -0.93] and list(p.box_y) == [-0.97, -0.97, 0.97, 0.97]
    assert list(p.reward_weight) == [1, 0, 5, 0, 10] and p.reward_ratio == 0.1
    assert (p.lidar_range, p.tolerant_time, p.rs_max_dist, p.env_collide, p.auto_reset) == (10.0, 200, 10.0, 0, 1)


def test_invalid_arguments_are_reported_not_crashed():
    from hope_b200 import capi
    lib = capi.load_library()
    assert lib.hope_create(None, 0, 16, 16, None) == -1
    ctx = C.c_void_p()
    assert lib.hope_create(C.byref(ctx), 0, 0, 16, None) == -1
    assert lib.hope_destroy(None) == -1
    assert lib.hope_generate_scenes(0, 0, 1, 1, None, None, None, None, None, None) == -1
    with pytest.raises(capi.HopeError):
        capi.check(-3)


def test_env_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from hope_b200 import capi
    from hope_b200.batched_env import BatchedParkingEnv
    with pytest.raises(capi.HopeError):
        BatchedParkingEnv(8)


@pytest.mark.parametrize("level", ["Normal", "Complex", "Extrem"])
def test_generated_scenes_are_valid_and_deterministic(level):
    """parking_map_normal.py:40-494 semantics: start/dest boxes touch nothing, gaps in budget,
    bounds = floor/ceil of the poses -/+ 10; same seed -> same scenes for any thread count."""
    from hope_b200.batched_env import generate_scenes
    from oracle import geom
    n = 200
    a = generate_scenes(n, level, 123, nthreads=1)
    b = generate_scenes(n, level, 123, nthreads=5)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    box = np.array([(-0.93, -0.97), (3.76, -0.97), (3.76, 0.97), (-0.93, 0.97)])

    def ring(pose):
        c, s = np.cos(pose[2]), np.sin(pose[2])
        pts = [(c * x - s * y + pose[0], s * x + c * y + pose[1]) for x, y in box]
        return pts + [pts[0]]

    nobs = (a["nverts"] > 0).sum(axis=1)
    assert 3 <= nobs.min() and nobs.max() <= 16
    if level == "Extrem":
        assert (a["case_id"] == 1).all()
    else:
        assert 0.3 < (a["case_id"] == 0).mean() < 0.7
    for i in range(n):
        srt, dst = ring(a["start"][i]), ring(a["dest"][i])
        assert not geom.rings_intersect(srt, dst)
        for k in range(16):
            nv = a["nverts"][i, k]
            if nv == 0:
                continue
            ob = [tuple(p) for p in a["obs"][i, k, :nv]]
            ob.append(ob[0])
            assert not geom.rings_intersect(srt, ob), (i, k)
            assert not geom.rings_intersect(dst, ob), (i, k)
        lo = np.minimum(a["start"][i, :2], a["dest"][i, :2]); hi = np.maximum(a["start"][i, :2], a["dest"][i, :2])
        assert np.array_equal(a["bounds"][i], [np.floor(lo[0] - 10), np.ceil(hi[0] + 10), np.floor(lo[1] - 10), np.ceil(hi[1] + 10)])


def test_product_tables_equal_reference_tables_bit_for_bit(golden_dir):
    import hashlib
    from hope_b200 import tables
    g = np.load(os.path.join(golden_dir, "mask_table.npz"))
    t = tables.host_tables()
    assert np.array_equal(t["mask_base"], g["vehicle_lidar_base"])
    assert np.array_equal(t["lidar_base"], g["vehicle_boundary"])
    assert np.array_equal(tables.discrete_actions(), g["discrete_actions"])
    assert hashlib.sha256(t["dist_star"].tobytes()).digest() == g["dist_star_sha256"].tobytes()
    theta = np.array([i * np.pi / 120 * 2 for i in range(120)])
    assert np.array_equal(t["ray_a"], np.sin(theta)) and np.array_equal(t["ray_b"], -np.cos(theta))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hope_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "parking_oracle" not in src, f
    code = "import sys; import hope_b200, hope_b200.batched_env, hope_b200.tables; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, check=True).stdout.strip()
    assert out == "False"


def test_host_mask_expansion_equals_the_reference_formula():
    """hope_step_host ships the action mask as uint8 step counts and rebuilds the float64 mask on the host
    (hope_expand_mask); action_mask.py:182-183: steps / 10, or 0.01 everywhere when the env has no free step."""
    from hope_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(3)
    steps = rng.integers(0, 11, size=(4096, 42), dtype=np.uint8)
    steps[::7] = 0                      # blocked envs: every action has 0 free steps
    steps[1::7, :] = np.minimum(steps[1::7, :], 1)
    mask = np.full((4096, 42), -1.0)
    capi.check(lib.hope_expand_mask(steps.ctypes.data, mask.ctypes.data, 4096))
    want = steps.astype(np.float64) / 10
    want[steps.sum(axis=1) == 0] = 0.01
    assert np.array_equal(mask, want)
    assert lib.hope_expand_mask(None, mask.ctypes.data, 1) == -1
    for n in (4096, 4095, 1, 0):        # the portable routine and ragged tails of the 8-wide one
        for fn in (lib.hope_expand_mask, lib.hope_expand_mask_portable):
            m2 = np.full((4096, 42), -1.0)
            capi.check(fn(steps.ctypes.data, m2.ctypes.data, n))
            assert np.array_equal(m2[:n], want[:n]) and (m2[n:] == -1.0).all()


def test_host_lidar_expansion_rebuilds_the_rows_bit_for_bit():
    """hope_step_host ships, per env, 120 flag bits + an offset + only the lidar values that differ from the per-ray no-hit
    constant lidar_range - lidar_base[ray] (lidar_simulator.py:46, 134); hope_expand_lidar rebuilds the float64 rows.  Packed
    here the way k_pack_lidar does (envs in arbitrary order in the value array), expanded by both routines."""
    from hope_b200 import capi, tables
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    n = 3000
    nohit = 10.0 - tables.host_tables()["lidar_base"]
    lidar = np.tile(nohit, (n, 1))
    hit = rng.random((n, 120)) < 0.57
    hit[::11] = False                    # nothing in range at all
    hit[1::11] = True                    # every beam hits
    hit[2::11, :] = False; hit[2::11, 119] = True; hit[2::11, 0] = True
    vals = rng.uniform(-1.0, 9.0, size=(n, 120))
    vals[5, 7] = -0.0; vals[6, 8] = np.nextafter(nohit[8], 0.0)    # a signed zero and a value one ulp off the constant are kept as they are
    lidar[hit] = vals[hit]
    keep = lidar.view(np.int64) != np.tile(nohit, (n, 1)).view(np.int64)
    bits = np.zeros((n, 4), dtype=np.uint32)
    for j in range(120):
        bits[:, j // 32] |= keep[:, j].astype(np.uint32) << np.uint32(j % 32)
    order = rng.permutation(n)           # atomics hand out the space in any order
    off = np.zeros(n, dtype=np.uint32)
    packed = np.zeros(int(keep.sum()) + 8)
    pos = 0
    for i in order:
        k = keep[i]
        off[i] = pos
        packed[pos:pos + k.sum()] = lidar[i, k]
        pos += int(k.sum())
    for portable in (0, 1):
        for m in (n, n - 1, 1, 0):
            out = np.full((n, 120), np.nan)
            capi.check(lib.hope_expand_lidar(bits.ctypes.data, off.ctypes.data, packed.ctypes.data, nohit.ctypes.data, out.ctypes.data, m, portable))
            assert np.array_equal(out[:m].view(np.int64), lidar[:m].view(np.int64)), portable
            assert np.isnan(out[m:]).all()
    assert lib.hope_expand_lidar(None, off.ctypes.data, packed.ctypes.data, nohit.ctypes.data, out.ctypes.data, 1, 0) == -1


def test_policy_weight_packing_follows_its_definition():
    """hope_policy_pack_matrix (host): float32 [out][in] -> bf16 (round to nearest even), input width zero-padded, stored in the
    order the tensor-core B fragments are read: packed[((nt * KS + ks) * 32 + lane) * 4 + 2 * half + e] =
    W[8 nt + lane / 4][16 ks + 8 half + 2 (lane % 4) + e]."""
    import torch
    from hope_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(2)
    for n_out, n_in, k_pad in ((128, 120, 128), (128, 5, 16), (768, 128, 128), (8, 42, 48)):
        w = rng.standard_normal((n_out, n_in)).astype(np.float32)
        w[0, 0] = 1.00390625          # exactly half way between two bf16 values: ties to even
        out = np.zeros(n_out * k_pad, dtype=np.uint16)
        capi.check(lib.hope_policy_pack_matrix(w.ctypes.data, n_out, n_in, k_pad, out.ctypes.data))
        ref = torch.zeros((n_out, k_pad), dtype=torch.bfloat16)
        ref[:, :n_in] = torch.from_numpy(w).to(torch.bfloat16)
        ref16 = ref.view(torch.int16).numpy().view(np.uint16)
        ks_total = k_pad // 16
        nt, ks, lane, half, e = np.meshgrid(np.arange(n_out // 8), np.arange(ks_total), np.arange(32), np.arange(2), np.arange(2), indexing="ij")
        want = ref16[8 * nt + lane // 4, 16 * ks + 8 * half + 2 * (lane % 4) + e].reshape(-1)
        assert np.array_equal(out, want), (n_out, n_in, k_pad)
    assert lib.hope_policy_pack_matrix(w.ctypes.data, 12, 5, 16, out.ctypes.data) == -1   # rows not a multiple of 8
