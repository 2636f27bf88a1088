"""TEST INFRASTRUCTURE ONLY — scenes that sit on the decision boundaries of the LiDAR raycast (shared by
tests/test_observe_host.py and oracle/make_adversarial_golden.py)."""
import numpy as np


def adversarial_scenes(rng, n):
    """Scenes built to sit on the decision boundaries of the raycast: vertices exactly on beam directions, edges collinear
    with a beam (det == 0), axis-aligned edges in the ego frame (degenerate bounding boxes, accepted only if the quotient
    lands exactly on them), vertices at exactly the 10 m range, slivers, and obstacles touching the ego position."""
    theta = np.array([i * np.pi / 120 * 2 for i in range(120)])
    start, obs, nverts = np.zeros((n, 3)), np.zeros((n, 16, 4, 2)), np.zeros((n, 16), dtype=np.int32)
    for i in range(n):
        kind = i % 6
        h = [0.0, np.pi / 2, np.pi, -np.pi / 2, float(theta[rng.integers(120)]), float(rng.uniform(-np.pi, np.pi))][kind]
        pos = rng.integers(-20, 20, size=2).astype(np.float64) * (0.5 if i % 2 else 1.0) if kind < 5 else rng.uniform(-20, 20, size=2)
        start[i] = [pos[0], pos[1], h]
        c, s = np.cos(h), np.sin(h)
        k = 0
        for _ in range(int(rng.integers(1, 9))):
            mode = int(rng.integers(0, 6))
            if mode == 0:    # quad with vertices exactly on four beam directions (ego frame)
                b = np.sort(rng.choice(120, size=4, replace=False))
                r = rng.uniform(1.0, 12.0, size=4)
                ego = np.stack([r * np.cos(theta[b]), r * np.sin(theta[b])], axis=1)
            elif mode == 1:  # axis-aligned box in the ego frame at integer / half-integer offsets
                x0, y0 = rng.integers(-9, 9, size=2) * 0.5
                w, hgt = rng.integers(1, 8, size=2) * 0.5
                ego = np.array([[x0, y0], [x0 + w, y0], [x0 + w, y0 + hgt], [x0, y0 + hgt]])
            elif mode == 2:  # triangle with one edge collinear with a beam through the ego position
                b = int(rng.integers(120)); r0, r1 = np.sort(rng.uniform(0.5, 11.0, size=2))
                d = np.array([np.cos(theta[b]), np.sin(theta[b])])
                ego = np.array([r0 * d, r1 * d, r0 * d + rng.uniform(-3, 3, size=2)])
            elif mode == 3:  # vertices at exactly the lidar range
                b = rng.choice(120, size=3, replace=False)
                ego = np.stack([10.0 * np.cos(theta[b]), 10.0 * np.sin(theta[b])], axis=1)
            elif mode == 4:  # sliver
                p = rng.uniform(-9, 9, size=2); d = rng.uniform(-4, 4, size=2)
                ego = np.array([p, p + d, p + d + 1e-9 * rng.uniform(-1, 1, size=2), p + 1e-9 * rng.uniform(-1, 1, size=2)])
            else:            # random quad, sometimes touching the ego position
                ego = rng.uniform(-11, 11, size=(4, 2))
                if rng.random() < 0.2:
                    ego[0] = 0.0
            world = np.stack([c * ego[:, 0] - s * ego[:, 1] + pos[0], s * ego[:, 0] + c * ego[:, 1] + pos[1]], axis=1)
            obs[i, k, :len(world)] = world
            nverts[i, k] = len(world)
            k += 1
    return start, obs, nverts
