"""The Reeds-Shepp word enumeration that k_rs_enumerate runs (hope_b200/csrc/rs_words.cuh) compiled with g++ and the
host libm (tests/rs_host_harness.cpp), replaying the 400 calc_all_paths known answers recorded from the unmodified
reference (tests/golden/reeds_shepp.npz).  Word count, insertion order, segment types and signs must be identical;
lengths are bit-identical except where the shared polar frames reorder a last-ulp rounding (a handful of words, a few
1e-15), and L is summed left to right (CPython 3.8 semantics; the recording ran on 3.12's compensated sum)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rs(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("rs_host") / "rs_host.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "host_stubs"),
                    "-o", out, os.path.join(ROOT, "tests", "rs_host_harness.cpp")], check=True, env=env)
    lib = C.CDLL(out)
    lib.rs_host_words.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rs_host_samples.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 4
    return lib


def test_known_answers_through_the_product_enumeration(rs, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "reeds_shepp.npz")))
    maxc = float(g["maxc"])
    exact_len = exact_L = words = 0
    for i in range(len(g["q"])):
        q = np.ascontiguousarray(g["q"][i], dtype=np.float64)
        cnt = C.c_int(0)
        nseg = np.zeros(16, dtype=np.int32); types = np.zeros((16, 5), dtype=np.uint8)
        lens = np.zeros((16, 5)); L = np.zeros(16)
        ctr = (C.c_ulonglong * 8)()
        assert rs.rs_host_words(q.ctypes.data, maxc, C.byref(cnt), nseg.ctypes.data, types.ctypes.data, lens.ctypes.data, L.ctypes.data, ctr) == 0
        k = int(g["npaths"][i])
        assert cnt.value == k, i
        assert np.array_equal(nseg[:k], g["nseg"][i, :k]) and np.array_equal(types[:k], g["types"][i, :k]), i
        assert np.array_equal(np.sign(lens[:k]), np.sign(g["lengths"][i, :k])), i
        np.testing.assert_allclose(lens[:k], g["lengths"][i, :k], rtol=0, atol=1e-13)
        np.testing.assert_allclose(L[:k], g["L"][i, :k], rtol=0, atol=1e-13)
        words += k
        exact_len += int((lens[:k] == g["lengths"][i, :k]).all(axis=1).sum())
        exact_L += int((L[:k] == g["L"][i, :k]).sum())
        assert ctr[3] == 0 and ctr[4] == 0   # no capacity overflow, no zero-length word
    assert words > 2000
    assert exact_len >= 0.99 * words   # bit-identical lengths for all but a handful of words
    assert exact_L >= 0.9 * words


def test_sample_chain_through_the_product_walker(rs, golden_dir):
    """generate_local_course + the map-frame transform (reeds_shepp.py:452-507, 46-49): sample count T, the first and
    last three map-frame samples and the coordinate sums of every known word, from k_rs_walk's plan and k_rs_check's
    per-lane replay compiled for the host.  The reference's trailing `while px[-1] == 0.0: pop` is applied here the
    way k_rs_check's zero-tail rule accounts for it."""
    g = dict(np.load(os.path.join(golden_dir, "reeds_shepp.npz")))
    maxc = float(g["maxc"])
    cap = 40000
    gx, gy, gyaw, lx = (np.zeros(cap) for _ in range(4))
    words = exact = long_words = 0
    for i in range(len(g["q"])):
        q = np.ascontiguousarray(g["q"][i], dtype=np.float64)
        for k in range(int(g["npaths"][i])):
            n = rs.rs_host_samples(q.ctypes.data, maxc, 0.1, k, cap, gx.ctypes.data, gy.ctypes.data, gyaw.ctypes.data, lx.ctypes.data)
            assert n > 0, (i, k, n)
            while n > 0 and lx[n - 1] == 0.0:   # reeds_shepp.py:500-505
                n -= 1
            assert n == int(g["T"][i, k]), (i, k)
            m = min(3, n)
            head = np.stack([gx[:m], gy[:m], gyaw[:m]], axis=1)
            tail = np.stack([gx[n - m:n][::-1], gy[n - m:n][::-1], gyaw[n - m:n][::-1]], axis=1)
            np.testing.assert_allclose(head, g["head"][i, k, :m], rtol=0, atol=1e-12)
            np.testing.assert_allclose(tail, g["tail"][i, k, :m], rtol=0, atol=1e-12)
            np.testing.assert_allclose([gx[:n].sum(), gy[:n].sum(), gyaw[:n].sum()], g["csum"][i, k], rtol=0, atol=1e-6)
            exact += int(np.array_equal(head, g["head"][i, k, :m]) and np.array_equal(tail, g["tail"][i, k, :m]))
            words += 1
            long_words += n > 256
    assert words > 2000 and long_words > 50     # words longer than one 256-sample chunk were walked on
    assert exact >= 0.95 * words
