"""Executable form of the exactness argument behind k_observe's action-mask sweep ("Mask sweep"):
a numpy model of the screened sweep the kernel runs, on the product's own tables, against the reference's literal
1 200 x 42 x 10 comparison (action_mask.py:145-184).  The kernel itself is checked on the GPU; this pins the claims
the screens rest on: (1) the first exceedance of dist_star[rho][j][:] equals the first exceedance of its running
maximum, (2) a beam whose two end distances both reach gpmax cannot lower any action through its 10 upsampled rays,
(3) a ray lowers action j below its current bound s only if d < P[rho][s-1][j]."""
import numpy as np

from hope_b200 import tables

NRAY, NUP, NACT, NITER = 120, 1200, 42, 10


def _literal(ds, w_lo, w_hi, L):
    """action_mask.py:145-184 up to the minimum over rays (before post_process)."""
    d = np.empty(NUP)
    for q in range(NRAY):
        qn = (q + 1) % NRAY
        d[10 * q:10 * q + 10] = L[q] * w_lo + L[qn] * w_hi     # linear upsample, :158-162
    exceed = ds > d[:, None, None]                                # (1200, 42, 10)
    first = np.where(exceed.any(axis=2), exceed.argmax(axis=2), NITER)
    return first.min(axis=0), d


def _screened(P, pmax, gpmax, w_lo, w_hi, L):
    """What k_observe does: beam screen, ray screen, bound screen, then the scan of the running maxima."""
    s = np.full(NACT, NITER)
    scans = rays = 0
    for q in range(NRAY):
        qn = (q + 1) % NRAY
        if not (min(L[q], L[qn]) * (1.0 - 1e-15) < gpmax[q]):
            continue
        for r in range(10):
            rho = 10 * q + r
            d = L[q] * w_lo[r] + L[qn] * w_hi[r]
            if not d < pmax[rho]:
                continue
            rays += 1
            need = (s > 0) & (d < P[rho, np.maximum(s, 1) - 1, np.arange(NACT)])
            if not need.any():
                continue
            scans += 1
            exceed = P[rho] > d                                   # (10, 42) running maxima, k major
            first = np.where(exceed.any(axis=0), exceed.argmax(axis=0), NITER)
            s = np.minimum(s, first)
    return s, rays, scans


def test_screened_sweep_equals_the_literal_sweep():
    tb = tables.host_tables()
    ds, w_lo, w_hi = tb["dist_star"], tb["w_lo"], tb["w_hi"]
    P = np.maximum.accumulate(ds, axis=2).transpose(0, 2, 1).copy()   # [rho][k][j], the kernel's pmaxk
    pmax = ds.max(axis=(1, 2))
    gpmax = pmax.reshape(NRAY, 10).max(axis=1)
    # claim (1) on the whole table, against arbitrary thresholds
    rng = np.random.default_rng(0)
    for d in rng.uniform(0, 12, size=6):
        a = np.where((ds > d).any(axis=2), (ds > d).argmax(axis=2), NITER)
        b = np.where((P > d).any(axis=1), (P > d).argmax(axis=1), NITER)
        assert np.array_equal(a, b)
    base = tb["mask_base"]
    total_rays = total_scans = lowered = 0
    for trial in range(60):
        kind = trial % 4
        if kind == 0:       # open space
            lidar = rng.uniform(6, 10, size=NRAY)
        elif kind == 1:     # a wall on one side
            lidar = np.full(NRAY, 10.0)
            k0 = int(rng.integers(0, NRAY)); w = int(rng.integers(5, 40))
            lidar[np.arange(k0, k0 + w) % NRAY] = rng.uniform(0.05, 3.0)
        elif kind == 2:     # clutter everywhere
            lidar = rng.uniform(0.0, 10.0, size=NRAY)
        else:               # boxed in
            lidar = rng.uniform(0.0, 0.6, size=NRAY)
        L = np.clip(lidar, 0, 10) + base                          # action_mask.py:170
        want, _ = _literal(ds, w_lo, w_hi, L)
        got, rays, scans = _screened(P, pmax, gpmax, w_lo, w_hi, L)
        assert np.array_equal(got, want), trial
        total_rays += rays; total_scans += scans; lowered += int((want < NITER).any())
    assert lowered >= 25 and 0 < total_scans < total_rays   # the bound screen does stop rays before the scan
