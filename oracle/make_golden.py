#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Runs in the build container only (needs /root/reference, which does not exist on the GPU
box).  The reference sources under /root/reference/src are imported as they lie, with
oracle/refshim in front of sys.path to stand in for the four third-party packages that are
not installed (shapely, gym, pygame, heapdict — see oracle/refshim/README.md).  Instrumentation
is by wrapping bound methods from outside; no reference file is edited or copied.

Fixtures written (all float64 unless noted):
  episodes_<Level>.npz   lock-step traces of CarParkingWrapper.step (env_wrapper.py:73-81):
                         scene, float64 actions, pose, status, lidar, mask, target, reward,
                         reward_info, RS result, substep counters        (BASELINE cfg 1)
  episodes_follow_<Level>.npz  same, but once an RS word is found its open-loop actions are
                         executed (covers ARRIVED, the box-union reward and long RS hand-offs)
  scenes_<Level>.npz     reference-generated scenes (parking_map_normal.py:460-494)
  reeds_shepp.npz        calc_all_paths known answers (reeds_shepp.py:35-54)
  images_<Level>.npz     CarParkingWrapper with use_img_observation=True: per step the uint8 image
                         (obs['img'] * 255, exact), pose, trajectory length, substep counters; pygame is the
                         software restatement oracle/softraster.py (unpinned), cv2 is the real one
  episodes_collide.npz   random-action episodes of all three levels recorded with ENV_COLLIDE = True (configs.py:79): a collision on
                         the first substep ends the episode with status COLLIDED (car_parking_base.py:264-267, 279-282)
  mask_table.npz         ActionMask constants: vehicle_lidar_base, sha256 + strided sample of
                         dist_star (action_mask.py:114-143), LidarSimlator.vehicle_boundary

Usage:  python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import hashlib
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_OBS = 16


def _import_reference(ref_root):
    src = os.path.join(ref_root, "src")
    sys.path.insert(0, src)
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    os.chdir(src)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import env.car_parking_base as cpb
    import env.env_wrapper as wrap
    import env.vehicle as vehicle
    import env.reeds_shepp as rs
    import env.parking_map_normal as pmn
    import configs
    return cpb, wrap, vehicle, rs, pmn, configs


def scene_arrays(m):
    """(start[3], dest[3], bounds[4], obs[MAX_OBS,4,2], nverts[MAX_OBS]) from a reference map."""
    obs = np.zeros((MAX_OBS, 4, 2))
    nverts = np.zeros(MAX_OBS, dtype=np.int32)
    assert len(m.obstacles) <= MAX_OBS
    for k, area in enumerate(m.obstacles):
        c = np.array(area.shape.coords)[:-1]
        assert c.shape[0] <= 4
        obs[k, :c.shape[0]] = c
        nverts[k] = c.shape[0]
    start = np.array(m.start.get_pos(), dtype=np.float64)
    dest = np.array(m.dest.get_pos(), dtype=np.float64)
    bounds = np.array([m.xmin, m.xmax, m.ymin, m.ymax], dtype=np.float64)
    return start, dest, bounds, obs, nverts


TYPE_CODE = {"S": 0, "L": 1, "R": 2}


def encode_path(path):
    types = np.full(5, 255, dtype=np.uint8)
    lens = np.zeros(5)
    for i, (c, l) in enumerate(zip(path.ctypes, path.lengths)):
        types[i] = TYPE_CODE[c]
        lens[i] = l
    return len(path.ctypes), types, lens


def plan_actions(types, lengths, step_ratio):
    """Open-loop unit actions for an RS word, the way the reference's planner consumes
    `info['path_to_dest']` (parking_agent.py:12-41): steer in {+1,0,-1}, signed length split
    into pieces of at most one env-step (step_ratio metres)."""
    steer_of = {1: 1.0, 0: 0.0, 2: -1.0}
    acts = []
    for t, l in zip(types, lengths):
        if t == 255:
            break
        rem = l / step_ratio
        sgn = 1.0 if rem > 0 else -1.0
        while abs(rem) > 1:
            acts.append([steer_of[int(t)], sgn])
            rem -= sgn
        if abs(rem) > 1e-3:
            acts.append([steer_of[int(t)], rem])
    return acts


def record_episodes(mods, level, n_episodes, seed, steps_per_episode=200, follow_rs=False, env_collide=False):
    cpb, wrap, vehicle, rs, pmn, configs = mods
    # ENV_COLLIDE (configs.py:79) is read by CarParking.step as a module global of car_parking_base (star-imported from configs):
    # setting it there is what editing configs.py does, without touching a reference file
    cpb.ENV_COLLIDE = bool(env_collide)
    raw = cpb.CarParking(render_mode="rgb_array", fps=100, verbose=False,
                         use_lidar_observation=True, use_img_observation=False, use_action_mask=True)
    env = wrap.CarParkingWrapper(raw)

    counters = {"substeps": 0, "retreats": 0, "valid_calls": [], "n_cand": 0}
    veh = raw.vehicle
    orig_step, orig_retreat = veh.step, veh.retreat

    def step_spy(action, step_time=configs.NUM_STEP):
        counters["substeps"] += 1
        return orig_step(action, step_time)

    def retreat_spy(prev):
        counters["retreats"] += 1
        return orig_retreat(prev)

    veh.step, veh.retreat = step_spy, retreat_spy
    orig_valid = raw.is_traj_valid

    def valid_spy(traj):
        r = orig_valid(traj)
        counters["valid_calls"].append((len(traj), bool(r)))
        return r

    raw.is_traj_valid = valid_spy
    orig_all = rs.calc_all_paths

    def all_spy(*a, **k):
        r = orig_all(*a, **k)
        counters["n_cand"] = len(r)
        return r

    rs.calc_all_paths = all_spy

    rec = {k: [] for k in (
        "ep", "action", "pose", "status", "lidar", "mask", "target", "reward", "reward_info", "done",
        "substeps", "retreated", "rs_found", "rs_nseg", "rs_types", "rs_lengths", "rs_L", "rs_ncand",
        "rs_ntried", "rs_T_last")}
    scn = {k: [] for k in ("start", "dest", "bounds", "obs", "nverts", "case_id", "reset_lidar",
                           "reset_mask", "reset_target")}
    for ep in range(n_episodes):
        np.random.seed(seed + ep)
        obs0 = env.reset(None, None, level)
        s, d, b, o, nv = scene_arrays(raw.map)
        scn["start"].append(s); scn["dest"].append(d); scn["bounds"].append(b)
        scn["obs"].append(o); scn["nverts"].append(nv); scn["case_id"].append(raw.map.case_id)
        scn["reset_lidar"].append(obs0["lidar"]); scn["reset_mask"].append(obs0["action_mask"])
        scn["reset_target"].append(obs0["target"])
        rng = np.random.default_rng(seed + 1000 * (ep + 1))
        queue = []
        for _ in range(steps_per_episode):
            a = rng.uniform(-1.0, 1.0, size=2)  # float64 (NumPy-2 promotion trap)
            if queue:
                a = np.array(queue.pop(0), dtype=np.float64)
            counters["substeps"] = 0; counters["retreats"] = 0
            counters["valid_calls"] = []; counters["n_cand"] = 0
            obs, reward, done, info = env.step(a)
            st = raw.vehicle.state
            rec["ep"].append(ep); rec["action"].append(a)
            rec["pose"].append([st.loc.x, st.loc.y, st.heading])
            rec["status"].append(info["status"].value)
            rec["lidar"].append(obs["lidar"]); rec["mask"].append(obs["action_mask"])
            rec["target"].append(obs["target"]); rec["reward"].append(reward)
            rec["reward_info"].append([float(v) for v in info["reward_info"].values()])
            rec["done"].append(done)
            rec["substeps"].append(counters["substeps"]); rec["retreated"].append(counters["retreats"])
            p = info["path_to_dest"]
            if p is not None:
                n, t, l = encode_path(p)
                rec["rs_found"].append(1); rec["rs_nseg"].append(n); rec["rs_types"].append(t)
                rec["rs_lengths"].append(l); rec["rs_L"].append(p.L)
            else:
                rec["rs_found"].append(0); rec["rs_nseg"].append(0)
                rec["rs_types"].append(np.full(5, 255, dtype=np.uint8)); rec["rs_lengths"].append(np.zeros(5))
                rec["rs_L"].append(0.0)
            rec["rs_ncand"].append(counters["n_cand"]); rec["rs_ntried"].append(len(counters["valid_calls"]))
            rec["rs_T_last"].append(counters["valid_calls"][-1][0] if counters["valid_calls"] else 0)
            if follow_rs and p is not None and not queue:
                queue = plan_actions(rec["rs_types"][-1], rec["rs_lengths"][-1], 1.25)
            if done:
                break
    rs.calc_all_paths = orig_all
    cpb.ENV_COLLIDE = False
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update({"scene_" + k: np.asarray(v) for k, v in scn.items()})
    return out


def record_images(mods, level, n_episodes, seed, steps_per_episode=40, traj_render_len=None, render_traj=True):
    """Image observation of the unmodified reference (car_parking_base.py:301-350, observation_processor.py).
    traj_render_len / render_traj: configs.py:86-88 TRAJ_RENDER_LEN (with the TRAJ_COLORS it implies) and :105 RENDER_TRAJ are read
    by CarParking._render as module globals of car_parking_base (star-imported from configs): setting them there is what editing
    configs.py does, without touching a reference file."""
    cpb, wrap, vehicle, rs, pmn, configs = mods
    saved = (cpb.TRAJ_RENDER_LEN, cpb.TRAJ_COLORS, cpb.RENDER_TRAJ)
    if traj_render_len is not None:
        cpb.TRAJ_RENDER_LEN = int(traj_render_len)
        cpb.TRAJ_COLORS = list(map(tuple, np.linspace(np.array(configs.TRAJ_COLOR_LOW), np.array(configs.TRAJ_COLOR_HIGH),
                                                      int(traj_render_len), endpoint=True, dtype=np.uint8)))  # configs.py:87-88
    cpb.RENDER_TRAJ = bool(render_traj)
    raw = cpb.CarParking(render_mode="rgb_array", fps=100, verbose=False,
                         use_lidar_observation=True, use_img_observation=True, use_action_mask=True)
    env = wrap.CarParkingWrapper(raw)
    counters = {"substeps": 0, "retreats": 0}
    veh = raw.vehicle
    orig_step, orig_retreat = veh.step, veh.retreat

    def step_spy(action, step_time=configs.NUM_STEP):
        counters["substeps"] += 1
        return orig_step(action, step_time)

    def retreat_spy(prev):
        counters["retreats"] += 1
        return orig_retreat(prev)

    veh.step, veh.retreat = step_spy, retreat_spy

    def u8(img):
        q = np.rint(img * 255.0)
        assert np.array_equal(q / 255.0, img)
        return q.astype(np.uint8)

    rec = {k: [] for k in ("ep", "action", "pose", "img", "traj_len", "substeps", "retreated", "done")}
    scn = {k: [] for k in ("start", "dest", "bounds", "obs", "nverts", "reset_img")}
    for ep in range(n_episodes):
        np.random.seed(seed + ep)
        obs0 = env.reset(None, None, level)
        s, d, b, o, nv = scene_arrays(raw.map)
        scn["start"].append(s); scn["dest"].append(d); scn["bounds"].append(b)
        scn["obs"].append(o); scn["nverts"].append(nv); scn["reset_img"].append(u8(obs0["img"]))
        rng = np.random.default_rng(seed + 1000 * (ep + 1))
        drift = rng.uniform(-1.0, 1.0, size=2)
        for k in range(steps_per_episode):
            a = rng.uniform(-1.0, 1.0, size=2)
            if ep % 2 == 1:  # every other episode keeps a persistent bias so the vehicle travels and turns
                a = np.clip(0.6 * drift + 0.4 * a, -1.0, 1.0)
                if k % 15 == 14:
                    drift = rng.uniform(-1.0, 1.0, size=2)
            counters["substeps"] = 0; counters["retreats"] = 0
            obs, reward, done, info = env.step(a)
            st = raw.vehicle.state
            rec["ep"].append(ep); rec["action"].append(a)
            rec["pose"].append([st.loc.x, st.loc.y, st.heading])
            rec["img"].append(u8(obs["img"])); rec["traj_len"].append(len(raw.vehicle.trajectory))
            rec["substeps"].append(counters["substeps"]); rec["retreated"].append(counters["retreats"])
            rec["done"].append(done)
            if done:
                break
    cpb.TRAJ_RENDER_LEN, cpb.TRAJ_COLORS, cpb.RENDER_TRAJ = saved
    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update({"scene_" + k: np.asarray(v) for k, v in scn.items()})
    return out


def record_scenes(mods, level, n, seed):
    cpb, wrap, vehicle, rs, pmn, configs = mods
    m = pmn.ParkingMapNormal(level)
    acc = {k: [] for k in ("start", "dest", "bounds", "obs", "nverts", "case_id")}
    for i in range(n):
        np.random.seed(seed + i)
        m.reset(None, None)
        s, d, b, o, nv = scene_arrays(m)
        acc["start"].append(s); acc["dest"].append(d); acc["bounds"].append(b)
        acc["obs"].append(o); acc["nverts"].append(nv); acc["case_id"].append(m.case_id)
    return {k: np.asarray(v) for k, v in acc.items()}


def record_reeds_shepp(mods, n, seed):
    cpb, wrap, vehicle, rs, pmn, configs = mods
    maxc = math.tan(configs.VALID_STEER[-1]) / configs.WHEEL_BASE
    rng = np.random.default_rng(seed)
    MAXP = 16
    q = np.zeros((n, 6))
    npaths = np.zeros(n, dtype=np.int32)
    nseg = np.zeros((n, MAXP), dtype=np.int32)
    types = np.full((n, MAXP, 5), 255, dtype=np.uint8)
    lens = np.zeros((n, MAXP, 5))
    Ls = np.zeros((n, MAXP))
    T = np.zeros((n, MAXP), dtype=np.int32)
    csum = np.zeros((n, MAXP, 3))
    head = np.zeros((n, MAXP, 3, 3))
    tail = np.zeros((n, MAXP, 3, 3))
    for i in range(n):
        r = rng.uniform(0.5, 12.0)
        th = rng.uniform(-math.pi, math.pi)
        sx, sy, syaw = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-2 * math.pi, 2 * math.pi)
        gx, gy = sx + r * math.cos(th), sy + r * math.sin(th)
        gyaw = rng.uniform(-2 * math.pi, 2 * math.pi)
        if i % 10 == 0:  # some axis-aligned / symmetric goals (tie and degenerate branches)
            syaw = 0.0; gyaw = [0.0, math.pi / 2, math.pi, -math.pi / 2][(i // 10) % 4]
            gx, gy = sx + round(r), sy + [0.0, 1.0, -2.0][(i // 40) % 3]
        q[i] = [sx, sy, syaw, gx, gy, gyaw]
        # the env passes an np.float64 heading (vehicle.py:93) and float x/y/goal; the scalar type
        # decides how builtin sum() accumulates word lengths (see oracle/c py_sum), so mimic it
        paths = rs.calc_all_paths(float(sx), float(sy), np.float64(syaw), float(gx), float(gy), float(gyaw), maxc, 0.1)
        assert len(paths) <= MAXP
        npaths[i] = len(paths)
        for k, p in enumerate(paths):
            nseg[i, k], types[i, k], lens[i, k] = encode_path(p)
            Ls[i, k] = p.L
            T[i, k] = len(p.x)
            csum[i, k] = [math.fsum(p.x), math.fsum(p.y), math.fsum(p.yaw)]
            for j in range(min(3, len(p.x))):
                head[i, k, j] = [p.x[j], p.y[j], p.yaw[j]]
                tail[i, k, j] = [p.x[-1 - j], p.y[-1 - j], p.yaw[-1 - j]]
    return dict(q=q, maxc=np.float64(maxc), npaths=npaths, nseg=nseg, types=types, lengths=lens, L=Ls,
                T=T, csum=csum, head=head, tail=tail)


def record_mask_table(mods):
    cpb, wrap, vehicle, rs, pmn, configs = mods
    import model.action_mask as am
    from env.lidar_simulator import LidarSimlator
    mask = am.ActionMask()
    lidar = LidarSimlator(configs.LIDAR_RANGE, configs.LIDAR_NUM)
    ds = np.ascontiguousarray(mask.dist_star)
    flat = ds.reshape(-1)
    return dict(vehicle_lidar_base=mask.vehicle_lidar_base, vehicle_boundary=lidar.vehicle_boundary,
                dist_star_shape=np.array(ds.shape), dist_star_sha256=np.frombuffer(
                    hashlib.sha256(ds.tobytes()).digest(), dtype=np.uint8),
                dist_star_stride=np.int64(97), dist_star_sample=flat[::97].copy(),
                vehicle_boxes=mask.vehicle_boxes,
                discrete_actions=np.array(configs.discrete_actions, dtype=np.float64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--scenes", type=int, default=96)
    ap.add_argument("--follow-episodes", type=int, default=12)
    ap.add_argument("--only", default=None, help="'rs', 'tables', 'images', 'trajcfg' or 'collide': regenerate just that fixture")
    ap.add_argument("--image-episodes", type=int, default=4)
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    mods = _import_reference(args.ref)
    os.makedirs(out, exist_ok=True)
    if args.only in (None, "tables"):
        np.savez_compressed(os.path.join(out, "mask_table.npz"), **record_mask_table(mods))
    if args.only in (None, "rs"):
        np.savez_compressed(os.path.join(out, "reeds_shepp.npz"), **record_reeds_shepp(mods, 400, 7))
    if args.only in (None, "images"):
        for level in ("Normal", "Complex", "Extrem"):
            im = record_images(mods, level, args.image_episodes, 777)
            np.savez_compressed(os.path.join(out, f"images_{level}.npz"), **im)
            print(level, "image steps", len(im["traj_len"]), "max traj", int(im["traj_len"].max()))
    if args.only in (None, "trajcfg"):
        # the image observation under other trajectory settings (what hope_set_render_traj mirrors)
        for tag, kw in (("len7", dict(traj_render_len=7)), ("off", dict(render_traj=False))):
            im = record_images(mods, "Normal", 2, 4242, steps_per_episode=30, **kw)
            im["traj_render_len"] = np.int32(kw.get("traj_render_len", 20) if kw.get("render_traj", True) else 0)
            np.savez_compressed(os.path.join(out, f"images_traj_{tag}.npz"), **im)
            print("trajectory settings", tag, "image steps", len(im["traj_len"]), "max traj", int(im["traj_len"].max()))
    if args.only in (None, "collide"):
        parts = [record_episodes(mods, level, 25, 9000 + 100 * k, env_collide=True) for k, level in enumerate(("Normal", "Complex", "Extrem"))]
        merged = {}
        ep_off = 0
        for part in parts:
            part = dict(part)
            part["ep"] = part["ep"] + ep_off
            ep_off += len(part["scene_start"])
            for k, v in part.items():
                merged.setdefault(k, []).append(v)
        merged = {k: np.concatenate(v) for k, v in merged.items()}
        np.savez_compressed(os.path.join(out, "episodes_collide.npz"), **merged)
        print("ENV_COLLIDE episodes", ep_off, "steps", len(merged["status"]), "status hist", np.bincount(merged["status"], minlength=6))
    if args.only is not None:
        return
    for level in ("Normal", "Complex", "Extrem"):
        np.savez_compressed(os.path.join(out, f"scenes_{level}.npz"), **record_scenes(mods, level, args.scenes, 42))
        ep = record_episodes(mods, level, args.episodes, 42)
        np.savez_compressed(os.path.join(out, f"episodes_{level}.npz"), **ep)
        fo = record_episodes(mods, level, args.follow_episodes, 4242, follow_rs=True)
        np.savez_compressed(os.path.join(out, f"episodes_follow_{level}.npz"), **fo)
        print(level, "follow steps", len(fo["status"]), "status hist", np.bincount(fo["status"], minlength=6))
        print(level, "steps", len(ep["status"]), "status hist", np.bincount(ep["status"], minlength=6),
              "rs found", int(ep["rs_found"].sum()), "retreats", int(ep["retreated"].sum()))


if __name__ == "__main__":
    main()
