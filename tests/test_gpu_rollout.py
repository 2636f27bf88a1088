"""GPU tests of the rollout-side rows (a19, f2): batched RS planner hand-off vs a restatement of
RsPlanner/ParkingAgent, masked discrete sampling, and the device-resident acting loop."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from hope_b200 import rollout  # noqa: E402
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes  # noqa: E402
from oracle import planner_oracle as plo  # noqa: E402


def test_planner_handoff_matches_rsplanner_semantics():
    """parking_agent.py:2-47, 60-110 and train_HOPE_sac.py:194-213, lock-step over 1 024 envs."""
    n, steps = 1024, 90
    sc = generate_scenes(2 * n, "Complex", 21)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=True)
    env.reset(); env.planner_reset()
    planners = [plo.PlannerOracle(1.25) for _ in range(n)]
    rng = np.random.default_rng(3)
    n_exec = n_loaded = n_finished = 0
    for t in range(steps):
        pol = rng.uniform(-1, 1, size=(n, 2))
        act, exe = env.planner_actions(torch.as_tensor(pol, device=env.device).contiguous())
        torch.cuda.synchronize()
        act, exe = act.cpu().numpy().copy(), exe.cpu().numpy().astype(bool).copy()
        want, want_exe = pol.copy(), np.zeros(n, dtype=bool)
        for i, p in enumerate(planners):
            if p.executing:
                want[i] = p.get_action(); want_exe[i] = True
                n_finished += not p.executing
        assert np.array_equal(exe, want_exe), t
        assert np.array_equal(act, want), t
        n_exec += int(exe.sum())
        env.step(torch.as_tensor(act, device=env.device).contiguous())
        torch.cuda.synchronize()
        o = {k: v.cpu().numpy() for k, v in env.out.items()}
        for i, p in enumerate(planners):
            if o["done"][i] or o["was_reset"][i]:
                p.reset()                                   # next episode: ParkingAgent.reset
            elif o["rs_found"][i]:
                k = int(o["rs_nseg"][i])
                before = p.executing
                p.set_path(o["rs_types"][i][:k], [float(v) for v in o["rs_lengths"][i][:k]])
                n_loaded += (not before)
    assert n_loaded > 20 and n_exec > 100 and n_finished > 5, (n_loaded, n_exec, n_finished)
    assert env.out["status"].eq(2).any() or True
    env.close()


def test_following_plans_reaches_the_slot():
    """Sanity of the whole hand-off: executing the found paths open-loop must produce ARRIVED episodes."""
    n = 2048
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", 5), auto_reset=True)
    env.reset(); env.planner_reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    arrived = 0
    for _ in range(150):
        pol = torch.rand((n, 2), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1
        act, _ = env.planner_actions(pol)
        env.step(act)
        arrived += int((env.out["status"] == 2).sum())
    assert arrived > 50, arrived
    env.close()


def test_masked_sampling_respects_mask_and_distribution():
    dev = torch.device("cuda")
    acts = rollout.possible_actions(dev)
    n = 20000
    mean = torch.tensor([[0.3, 0.6]], dtype=torch.float64, device=dev).expand(n, 2)
    std = torch.tensor([[0.5, 0.9]], dtype=torch.float64, device=dev).expand(n, 2)
    mask = torch.zeros((n, 42), dtype=torch.float64, device=dev)
    mask[:, 3:12] = torch.linspace(0.1, 1.0, 9, dtype=torch.float64, device=dev)
    mask[:, 25] = 0.5
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    a, idx = rollout.masked_discrete_actions(mean, std, mask, acts, gen)
    want = plo.masked_action_probabilities(mean[0].cpu().numpy(), std[0].cpu().numpy(), mask[0].cpu().numpy(), acts.cpu().numpy())
    got = np.bincount(idx.cpu().numpy(), minlength=42) / n
    assert (got[want == 0] == 0).all()
    assert np.abs(got - want).max() < 0.02
    assert torch.equal(a, acts[idx])


def test_rollout_engine_runs_device_resident():
    n = 4096
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", 9), auto_reset=True)
    actor = rollout.ReferenceShapedActor().to(env.device)
    eng = rollout.RolloutEngine(env, actor, seed=0)
    seen = {"steps": 0, "exec": 0}

    def store(t, obs, action, reward, done, log_prob, executing):
        assert action.shape == (n, 2) and action.dtype == torch.float64 and reward.shape == (n,)
        assert torch.isfinite(log_prob).all() and torch.isfinite(obs["lidar"]).all()
        seen["steps"] += 1
        seen["exec"] += int(executing.sum())

    c0 = env.counters()
    eng.collect(24, store=store)
    torch.cuda.synchronize()
    c1 = env.counters()
    assert seen["steps"] == 24
    assert c1["env_steps"] - c0["env_steps"] > 20 * n
    assert eng.norm.n == 24 * n and torch.isfinite(eng.norm.mean["lidar"]).all()
    env.close()


def test_stored_transition_pairs_the_action_with_the_observation_it_was_chosen_from():
    """the env writes its outputs in place: what `store` receives must be the observation act() saw, not the next one, and the
    log-probability must be that of the EXECUTED action (the plan's where a route is being executed, parking_agent.py:93-97)"""
    import math
    n = 2048
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", 4), auto_reset=True)
    actor = rollout.ReferenceShapedActor().to(env.device)
    eng = rollout.RolloutEngine(env, actor, seed=0)
    seen, dists = [], []
    orig_act = eng.act

    def spy_act(obs):
        seen.append({k: obs[k].clone() for k in ("lidar", "target", "action_mask")})
        a, dist = orig_act(obs)
        dists.append(dist)
        return a, dist

    eng.act = spy_act
    n_exec = [0]

    def store(t, obs, action, reward, done, log_prob, executing):
        for k in ("lidar", "target", "action_mask"):
            assert torch.equal(obs[k], seen[t][k]), (t, k)
        mean, std = dists[t]
        want = -0.5 * ((action - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
        assert torch.equal(log_prob, want)
        ex = executing.bool()
        n_exec[0] += int(ex.sum())
        if ex.any():  # plan actions are {-1,0,1} steers with unit or remainder speeds: not what the sampler drew
            assert torch.all((action[ex, 0] == 0) | (action[ex, 0].abs() == 1))

    eng.collect(40, store=store)
    torch.cuda.synchronize()
    assert len(seen) == 40 and n_exec[0] > 0
    assert not torch.equal(seen[0]["lidar"], seen[1]["lidar"])
    env.close()


def test_fused_state_norm_kernel_equals_the_torch_reference_and_the_recording(golden_dir):
    """hope_state_norm (csrc/policy_glue.cu) against RunningNorm (plain PyTorch float64, itself pinned on the reference's StateNorm by
    tests/test_rollout_helpers.py) and against the recording of the unmodified StateNorm directly"""
    import os
    g = dict(np.load(os.path.join(golden_dir, "f2_helpers.npz")))
    dev = torch.device("cuda")
    n_all = len(g["norm_lidar"])
    for batch in (96, 512, n_all):
        fused = rollout.FusedStateNorm(batch, dev)
        ref = rollout.RunningNorm({"lidar": (120,), "target": (5,)}, dev)
        for lo in range(0, n_all - batch + 1, batch):
            obs = {"lidar": torch.as_tensor(g["norm_lidar"][lo:lo + batch], device=dev).contiguous(),
                   "target": torch.as_tensor(g["norm_target"][lo:lo + batch], device=dev).contiguous(),
                   "action_mask": torch.rand((batch, 42), dtype=torch.float64, device=dev)}
            out = fused(obs)
            ref.update({"lidar": obs["lidar"], "target": obs["target"]})
            want = ref({"lidar": obs["lidar"], "target": obs["target"]})
            for k in ("lidar", "target"):
                assert out[k].dtype == torch.float32
                # float32 outputs: equal to the float64 reference within float32 rounding of values up to ~1e2
                assert torch.allclose(out[k].double(), want[k], rtol=2e-7, atol=2e-6), (batch, lo, k, (out[k].double() - want[k]).abs().max())
            assert torch.equal(out["action_mask"], obs["action_mask"].float())
        assert fused.n == ref.n
        for k in ("lidar", "target"):
            assert torch.allclose(fused.mean[k], ref.mean[k], rtol=1e-12, atol=1e-12)
            assert torch.allclose(fused.m2[k], ref.m2[k], rtol=1e-10, atol=1e-9)
        if fused.n == n_all:  # every recorded observation went in: the reference's final statistics
            std = torch.sqrt(fused.m2["lidar"] / fused.n).cpu().numpy()
            np.testing.assert_allclose(fused.mean["lidar"].cpu().numpy(), g["norm_mean_lidar"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(std, g["norm_std_lidar"], rtol=1e-10, atol=1e-12)
            probe = {"lidar": torch.as_tensor(g["norm_probe_lidar"], device=dev)[None].repeat(batch, 1).contiguous(),
                     "target": torch.as_tensor(g["norm_probe_target"], device=dev)[None].repeat(batch, 1).contiguous(),
                     "action_mask": torch.ones((batch, 42), dtype=torch.float64, device=dev)}
            out = fused(probe, update=False)
            np.testing.assert_allclose(out["lidar"][0].cpu().numpy(), g["norm_probe_out_lidar"], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(out["target"][3].cpu().numpy(), g["norm_probe_out_target"], rtol=1e-6, atol=1e-6)
            assert fused.n == n_all  # update=False leaves the statistics alone


def test_fused_masked_sampler_kernel_draws_from_choose_actions_distribution(golden_dir):
    """hope_masked_sample against the probabilities ActionMask.choose_action hands to np.random.choice (recorded from the reference)
    and against the inverse-CDF rule evaluated in PyTorch float64 with the kernel's own uniform variates"""
    import os
    g = dict(np.load(os.path.join(golden_dir, "f2_helpers.npz")))
    dev = torch.device("cuda")
    acts = rollout.possible_actions(dev)
    # (1) exact: given u, the drawn index is the first one whose cumulative weight exceeds u * total
    n = 50000
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    mean = (torch.rand((n, 2), device=dev, generator=gen) * 2.4 - 1.2).float().contiguous()   # partly outside [-1, 1]: clamped like ppo_agent.py:141
    log_std = torch.tensor([-0.4, 0.2], dtype=torch.float64, device=dev)
    mask = torch.round(torch.rand((n, 42), dtype=torch.float64, device=dev, generator=gen) * 10) / 10
    mask *= (torch.rand((n, 42), device=dev, generator=gen) < 0.5)
    mask[:, 5] = torch.clamp(mask[:, 5], min=0.1)
    smp = rollout.FusedMaskedSampler(n, dev, seed=11)
    action, idx = smp(mean, log_std, mask)
    action, idx = action.clone(), idx.clone()   # the sampler reuses its output buffers
    torch.cuda.synchronize()
    m = torch.clamp(mean.double(), -1, 1)
    std = torch.exp(log_std).expand_as(m)
    z = (acts.unsqueeze(0) - m.unsqueeze(1)) / std.unsqueeze(1)
    lp = -0.5 * z * z - torch.log(math.sqrt(2 * math.pi) * std).unsqueeze(1)
    e = torch.exp(torch.clamp(lp, -10, 10).sum(dim=2)) * mask
    cdf = torch.cumsum(e, dim=1)
    want = (cdf > (smp.u * cdf[:, -1]).unsqueeze(1)).float().argmax(dim=1)
    same = (want == idx.long())
    assert same.float().mean() > 0.9995, float(same.float().mean())      # sequential vs pairwise summation may flip a boundary case
    assert (mask.gather(1, idx.long().unsqueeze(1)) > 0).all()            # never a masked action
    assert torch.equal(action, acts[idx.long()])
    u = smp.u.cpu().numpy()
    assert abs(u.mean() - 0.5) < 0.01 and abs(np.quantile(u, 0.1) - 0.1) < 0.01 and 0 <= u.min() and u.max() < 1
    a2, idx2 = smp(mean, log_std, mask)                                   # next step: a different stream position
    assert (idx2 != idx).float().mean() > 0.3
    # (2) distribution: 20 000 draws per recorded (mean, std, mask) triple against the reference's probabilities
    for row in (0, 7, 50, 123):
        k = 20000
        mean_r = torch.as_tensor(g["choose_mean"][row], device=dev).float().expand(k, 2).contiguous()
        ls = torch.log(torch.as_tensor(g["choose_std"][row], device=dev))
        mask_r = torch.as_tensor(g["choose_mask"][row], device=dev).expand(k, 42).contiguous()
        s2 = rollout.FusedMaskedSampler(k, dev, seed=row)
        _, ix = s2(mean_r, ls, mask_r)
        got = np.bincount(ix.cpu().numpy(), minlength=42) / k
        want_p = g["choose_prob"][row]
        assert (got[want_p == 0] == 0).all()
        assert np.abs(got - want_p).max() < 4 * np.sqrt(0.25 / k) + 1e-3, (row, np.abs(got - want_p).max())


def test_rollout_engine_fused_path_and_eager_path_agree_in_distribution():
    """the engine with the CUDA glue kernels + graph-replayed policy vs the eager PyTorch engine: same running statistics after the
    same env steps when both are driven by the same actions"""
    n = 4096
    sc = generate_scenes(2 * n, "mix", 9)
    env_a = BatchedParkingEnv(n, scenes=sc, auto_reset=True)
    env_b = BatchedParkingEnv(n, scenes=sc, auto_reset=True)
    torch.manual_seed(0)
    actor = rollout.ReferenceShapedActor().to(env_a.device)
    fused = rollout.RolloutEngine(env_a, actor, seed=0, fused=True, graph=True, policy_kernel=False)
    eager = rollout.RolloutEngine(env_b, actor, seed=0, fused=False)
    assert "policy_glue" in fused.glue and "CUDA graph" in fused.glue and eager.glue == "eager PyTorch"
    for t in range(6):
        a, (mean_f, _) = fused.act(fused.obs)
        _, (mean_e, _) = eager.act(eager.obs)
        assert torch.allclose(mean_f, mean_e, atol=3e-2), (t, (mean_f - mean_e).abs().max())   # bf16 policy on float32-rounded vs float64-normalised inputs
        a = a.clone()
        fused.obs = env_a.step(a)[0]
        eager.obs = env_b.step(a)[0]
    assert fused.norm.n == eager.norm.n == 6 * n
    assert torch.allclose(fused.norm.mean["lidar"], eager.norm.mean["lidar"], rtol=1e-12, atol=1e-12)
    assert torch.allclose(fused.norm.m2["target"], eager.norm.m2["target"], rtol=1e-10, atol=1e-9)
    fused.collect(8)
    torch.cuda.synchronize()
    env_a.close(); env_b.close()


@pytest.mark.parametrize("n", [65, 4096])
def test_policy_forward_kernel_matches_the_float32_module(n):
    """hope_policy_forward (one kernel, bf16 operands, float32 accumulation) against the plain float32 PyTorch forward of the
    same 3-modal actor, default initialisation and a 3x scaled copy (larger pre-activations, saturating tanh).  Bar: no further
    from float32 than twice what PyTorch's own bf16 autocast forward (the path it replaces) is, plus 2e-3; and absolutely
    max |diff| <= 3e-2 on outputs in [-1, 1] at default initialisation."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    for scale in (1.0, 3.0):
        net = rollout.ReferenceShapedActor().to(dev).eval()
        with torch.no_grad():
            for name, p in net.named_parameters():
                if name.endswith("weight") and p.dim() == 2:
                    p.mul_(scale)
                if name.endswith("bias") or "norm" in name:
                    p.add_(0.1 * torch.randn_like(p))
        obs = {"lidar": torch.randn(n, 120, device=dev), "target": torch.randn(n, 5, device=dev), "action_mask": torch.rand(n, 42, device=dev)}
        obs["action_mask"][::5] = 0.0
        with torch.no_grad():
            ref = net(obs)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                low = net(obs).float()
        fp = rollout.FusedPolicy(net, n, dev)
        got = fp(obs).clone()
        torch.cuda.synchronize()
        assert got.shape == (n, 2) and torch.isfinite(got).all() and got.abs().max() <= 1.0
        err, err_autocast = (got - ref).abs(), (low - ref).abs()
        print(f"  scale {scale}: kernel max {err.max():.2e} mean {err.mean():.2e} | autocast max {err_autocast.max():.2e} mean {err_autocast.mean():.2e}")
        assert err.max() <= 2 * err_autocast.max() + 2e-3 and err.mean() <= 2 * err_autocast.mean() + 5e-4
        assert err.max() <= (3e-2 if scale == 1.0 else 1.5e-1)
        # a parameter change is picked up by refresh()
        with torch.no_grad():
            net.net.output[2].bias.add_(0.25)
            ref2 = net(obs)
        fp.refresh()
        assert (fp(obs) - ref2).abs().max() <= (3e-2 if scale == 1.0 else 1.5e-1) and (ref2 - ref).abs().max() > 0.05


def test_rollout_engine_uses_the_policy_kernel():
    n = 2048
    env = BatchedParkingEnv(n, scenes=generate_scenes(n, "Normal", 4), auto_reset=True)
    policy, _ = rollout.reference_actor(device=env.device)
    eng = rollout.RolloutEngine(env, policy, seed=3)
    assert eng.policy_kernel is not None and "hope_policy_forward" in eng.glue
    ref = rollout.RolloutEngine(BatchedParkingEnv(n, scenes=generate_scenes(n, "Normal", 4), auto_reset=True), policy, seed=3, policy_kernel=False)
    a, (mean_a, _) = eng.act(eng.obs)
    b, (mean_b, _) = ref.act(ref.obs)
    torch.cuda.synchronize()
    assert (mean_a - mean_b).abs().max() < 2e-2            # same observations, same statistics: kernel vs autocast graph forward
    assert (a == b).all(dim=1).float().mean() > 0.97       # same Philox stream: the draws differ only where the means' difference moves a boundary
    eng.collect(4)
    torch.cuda.synchronize()
    eng.env.close(); ref.env.close()


def test_overlapped_collect_equals_the_plain_loop():
    """RolloutEngine.collect without a store computes the next action next to the Reeds-Shepp kernels of the current step
    (hope_wait_observed + a second stream); kernels, inputs and per-stream order are the plain loop's, so after 12 steps the two
    engines must hold identical states, observations and running statistics."""
    n = 4096
    sc = generate_scenes(2 * n, "mix", 21)
    torch.manual_seed(5)
    actor = rollout.ReferenceShapedActor().to(torch.device("cuda", 0))
    engines = []
    for overlap in (True, False):
        env = BatchedParkingEnv(n, scenes=sc, auto_reset=True)
        eng = rollout.RolloutEngine(env, actor, seed=9, overlap=overlap)
        eng.collect(5); eng.collect(7)
        torch.cuda.synchronize()
        engines.append(eng)
    a, b = engines
    sa, sb = a.env.get_state(), b.env.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    for k in ("lidar", "target", "action_mask"):
        assert torch.equal(a.obs[k], b.obs[k]), k
    assert torch.equal(a.norm.stats, b.norm.stats) and a.norm.n == b.norm.n
    assert a.env.wait_observed(torch.cuda.current_stream()) is True
    a.env.close(); b.env.close()


@pytest.mark.parametrize("n", [3, 1000])
def test_img_conv_kernel_matches_the_float32_conv_stack(n):
    """hope_img_conv_forward (both residual conv blocks of the image encoder in one kernel, float32 arithmetic, tanh.approx,
    bf16 output) against PyTorch's float32 embed_img.net[0:3] on uint8 images: |diff| <= 2^-8 relative (the bf16 output
    rounding) + 4e-3 absolute (tanh.approx: 2^-11 relative per block)."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    net = rollout.ReferenceShapedActor(use_img=True).to(dev).eval()
    with torch.no_grad():
        for name, p in net.embed_img.named_parameters():
            if "net.0" in name or "net.1" in name:
                p.mul_(2.0).add_(0.05 * torch.randn_like(p))
    img = torch.randint(0, 256, (n, 3, 64, 64), dtype=torch.uint8, device=dev)
    img[0, :, :20] = 0; img[-1, :, :, 40:] = 255          # flat regions: borders and pooling ties
    with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
        want = net.embed_img.net[:3](img.float() / 255.0)
    conv = rollout.FusedImgConv(net, n, dev)
    got = conv(img).float()
    torch.cuda.synchronize()
    assert got.shape == want.shape == (n, 2048)
    err = (got - want).abs()
    print(f"  conv stack: max |diff| {err.max():.2e}, mean {err.mean():.2e}, |want| max {want.abs().max():.2f}")
    assert (err <= want.abs() * 2.0 ** -8 + 4e-3).all()
    with torch.no_grad():                                  # refresh() picks up a parameter change
        net.embed_img.net[1].shortcut[0].bias.add_(0.5)
        want2 = net.embed_img.net[:3](img.float() / 255.0)
    conv.refresh()
    assert ((conv(img).float() - want2).abs() <= want2.abs() * 2.0 ** -8 + 4e-3).all() and (want2 - want).abs().min() > 0.4


def test_rollout_engine_with_images_uses_the_conv_kernel():
    n = 1024
    sc = generate_scenes(n, "Normal", 6)
    torch.manual_seed(2)
    policy = rollout.ReferenceShapedActor(use_img=True).to(torch.device("cuda", 0)).eval()
    eng = rollout.RolloutEngine(BatchedParkingEnv(n, scenes=sc, auto_reset=True, use_img_observation=True), policy, seed=3)
    assert eng.img_conv is not None and eng.policy_kernel is not None and eng.policy_kernel.n_modal == 4
    assert "hope_img_conv_forward" in eng.glue and "hope_policy_forward" in eng.glue
    ref = rollout.RolloutEngine(BatchedParkingEnv(n, scenes=sc, auto_reset=True, use_img_observation=True), policy, seed=3, policy_kernel=False)
    assert ref.img_conv is None
    a, (mean_a, _) = eng.act(eng.obs)
    b, (mean_b, _) = ref.act(ref.obs)
    torch.cuda.synchronize()
    assert (mean_a - mean_b).abs().max() < 2e-2            # same observations and statistics: conv kernel vs cuDNN under autocast
    assert (a == b).all(dim=1).float().mean() > 0.97
    eng.collect(4)
    torch.cuda.synchronize()
    eng.env.close(); ref.env.close()


@pytest.mark.parametrize("n", [37, 2048])
def test_policy_forward_kernel_4_modal_matches_the_float32_module(n):
    """hope_policy_forward_img (the 4-token network: lidar, target, action mask, image token from the encoder's mean head) against
    the plain float32 PyTorch forward of the same module, image encoder included (conv kernel -> library GEMMs -> kernel vs
    cuDNN float32).  Same bar as the 3-modal kernel: no further from float32 than twice PyTorch's bf16 autocast forward + 2e-3."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(13)
    net = rollout.ReferenceShapedActor(use_img=True).to(dev).eval()
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("bias") or "norm" in name:
                p.add_(0.1 * torch.randn_like(p))
    assert rollout.FusedPolicy.supports(net) == 4
    obs = {"lidar": torch.randn(n, 120, device=dev), "target": torch.randn(n, 5, device=dev), "action_mask": torch.rand(n, 42, device=dev)}
    img = torch.randint(0, 256, (n, 3, 64, 64), dtype=torch.uint8, device=dev)
    obs["img"] = img.float() / 255.0
    with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
        ref = net(obs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            low = net(obs).float()
    conv, fp = rollout.FusedImgConv(net, n, dev), rollout.FusedPolicy(net, n, dev)
    assert fp.n_modal == 4
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        mean = net.img_mean_from_conv(conv(img)).float().contiguous()
    got = fp({**obs, "img_mean": mean}).clone()
    torch.cuda.synchronize()
    assert got.shape == (n, 2) and torch.isfinite(got).all() and got.abs().max() <= 1.0
    err, err_autocast = (got - ref).abs(), (low - ref).abs()
    print(f"  4-modal: kernel max {err.max():.2e} mean {err.mean():.2e} | autocast max {err_autocast.max():.2e} mean {err_autocast.mean():.2e}")
    assert err.max() <= 2 * err_autocast.max() + 2e-3 and err.mean() <= 2 * err_autocast.mean() + 5e-4
    assert err.max() <= 3e-2
    # the image token matters (the kernel does not ignore it), and refresh() picks up re_embed_img
    got0 = fp({**obs, "img_mean": torch.zeros_like(mean)}).clone()
    assert (got0 - got).abs().max() > 1e-3
    with torch.no_grad():
        net.re_embed_img[1].bias.add_(0.5)
        ref2 = net(obs)
    fp.refresh()
    assert (fp({**obs, "img_mean": mean}) - ref2).abs().max() <= 3e-2 and (ref2 - ref).abs().max() > 1e-3
