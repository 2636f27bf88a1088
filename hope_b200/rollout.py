"""GPU-resident rollout loop around BatchedParkingEnv — the batched counterpart of the reference's
episode loop (src/train/train_HOPE_ppo.py:190-219, train_HOPE_sac.py:191-225, scope row a19) with the
policy-side helpers that are batch-1 host code in the reference (row f2):

  masked_discrete_actions   ActionMask.choose_action     model/action_mask.py:199-227
  env.planner_actions       RsPlanner / ParkingAgent     model/agent/parking_agent.py:2-47, 60-110
  RunningNorm               StateNorm (Welford)          model/state_norm.py:25-46

The policy itself is the reference's (`model.network.MultiObsEmbedding` works with batch > 1); for
synthetic benchmarks without the reference tree, `ReferenceShapedActor` has the same layer shapes
(ACTOR_CONFIGS, configs.py:134-152, three modalities when the image is off).  Policy GEMMs run through
stock PyTorch (bf16 autocast -> tensor cores); everything else stays in float64 device tensors.
"""
import math

import torch
from torch import nn

from . import tables

N_ACTION = 42


def possible_actions(device, dtype=torch.float64):
    """configs.py:108-115 scaled like action_mask.py:218-222: steer / 0.75, speed / 1."""
    a = torch.as_tensor(tables.discrete_actions(), device=device, dtype=dtype)
    return a / torch.tensor([float(tables.refconfig.load().VALID_STEER[-1]), 1.0], device=device, dtype=dtype)


def masked_action_probs(mean, std, mask, actions):
    """mean, std: (N,2); mask: (N,42); actions: (42,2) -> p (N,42)   (action_mask.py:213-225)"""
    z = (actions.unsqueeze(0) - mean.unsqueeze(1)) / std.unsqueeze(1)
    logp = -0.5 * z * z - torch.log(math.sqrt(2 * math.pi) * std).unsqueeze(1)
    e = torch.exp(torch.clamp(logp, -10, 10).sum(dim=2)) * mask
    return e / e.sum(dim=1, keepdim=True)


def masked_discrete_actions(mean, std, mask, actions, generator=None):
    """Batched ActionMask.choose_action: sample one of the 42 discrete actions per env."""
    p = masked_action_probs(mean, std, mask, actions)
    idx = torch.multinomial(p, 1, generator=generator).squeeze(1)
    return actions[idx], idx


class RunningNorm(object):
    """StateNorm (model/state_norm.py:25-46): running mean/std of `lidar` and `target`, updated with a
    whole batch per step (parallel Welford merge), applied as (x - mean) / (std + 1e-8)."""

    def __init__(self, shapes, device):
        self.n = 0
        self.mean = {k: torch.zeros(s, dtype=torch.float64, device=device) for k, s in shapes.items()}
        self.m2 = {k: torch.zeros(s, dtype=torch.float64, device=device) for k, s in shapes.items()}

    def update(self, obs):
        b = next(iter(obs.values())).shape[0]
        tot = self.n + b
        for k in self.mean:
            x = obs[k]
            bm = x.mean(dim=0)
            delta = bm - self.mean[k]
            self.m2[k] += ((x - bm) ** 2).sum(dim=0) + delta * delta * (self.n * b / tot)
            self.mean[k] += delta * (b / tot)
        self.n = tot

    def __call__(self, obs):
        out = dict(obs)
        for k in self.mean:
            std = torch.sqrt(self.m2[k] / max(self.n, 1))
            out[k] = (obs[k] - self.mean[k]) / (std + 1e-8)
        return out


class ReferenceShapedActor(nn.Module):
    """Same shapes as MultiObsEmbedding(ACTOR_CONFIGS) with the image modality off: three 2-layer tanh
    embeddings -> 3 x 128 tokens -> one pre-norm transformer block (8 heads x 32, FF 128) ->
    Linear(384,128) tanh Linear(128,2) tanh   (network.py:34-196, attention.py:16-92)."""

    def __init__(self, lidar=120, target=5, mask=42, embed=128, heads=8, dim_head=32, mlp=128, hidden=128, out=2):
        super().__init__()
        emb = lambda d: nn.Sequential(nn.Linear(d, embed), nn.Tanh(), nn.Linear(embed, embed))
        self.embed_lidar, self.embed_tgt, self.embed_am = emb(lidar), emb(target), emb(mask)
        self.norm1, self.norm2 = nn.LayerNorm(embed), nn.LayerNorm(embed)
        self.heads, self.dim_head = heads, dim_head
        self.to_qkv = nn.Linear(embed, heads * dim_head * 3, bias=False)
        self.to_out = nn.Linear(heads * dim_head, embed)
        self.ff = nn.Sequential(nn.Linear(embed, mlp), nn.Tanh(), nn.Linear(mlp, embed))
        self.head = nn.Sequential(nn.Linear(3 * embed, hidden), nn.Tanh(), nn.Linear(hidden, out))
        self.log_std = nn.Parameter(torch.zeros(out))

    def forward(self, obs):
        x = torch.stack([self.embed_lidar(obs["lidar"]), self.embed_tgt(obs["target"]), self.embed_am(obs["action_mask"])], dim=1)
        b, n, _ = x.shape
        q, k, v = self.to_qkv(self.norm1(x)).view(b, n, 3, self.heads, self.dim_head).permute(2, 0, 3, 1, 4)
        # 3 tokens per env: explicit softmax(q k^T / sqrt(d)) v like attention.py:33-46 (the fused SDPA back ends
        # reject a 65 536 x 8-head batch of 3-token sequences)
        a = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (self.dim_head ** -0.5), dim=-1), v)
        x = self.to_out(a.transpose(1, 2).reshape(b, n, self.heads * self.dim_head)) + x
        x = self.ff(self.norm2(x)) + x
        return torch.tanh(self.head(x.reshape(b, n * x.shape[-1])))


class RolloutEngine(object):
    """PPO-style acting loop for N envs: normalise -> policy (bf16 autocast) -> masked discrete sampling ->
    RS plan override -> env.step, all on the device; no host synchronisation inside `collect`."""

    def __init__(self, env, policy, log_std=None, use_planner=True, use_mask_sampling=True, state_norm=True,
                 autocast_dtype=torch.bfloat16, seed=0):
        self.env, self.policy = env, policy
        dev = env.device
        self.actions42 = possible_actions(dev)
        self.use_planner, self.use_mask_sampling = use_planner, use_mask_sampling
        self.norm = RunningNorm({"lidar": (120,), "target": (5,)}, dev) if state_norm else None
        self.log_std = log_std if log_std is not None else getattr(policy, "log_std", torch.zeros(2, device=dev))
        self.autocast_dtype = autocast_dtype
        self.gen = torch.Generator(device=dev); self.gen.manual_seed(seed)
        self.obs = env.reset()
        if use_planner:
            env.planner_reset()

    @torch.no_grad()
    def act(self, obs):
        o = {"lidar": obs["lidar"], "target": obs["target"]}
        if self.norm is not None:
            self.norm.update(o)
            o = self.norm(o)
        net_in = {"lidar": o["lidar"].float(), "target": o["target"].float(), "action_mask": obs["action_mask"].float()}
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            mean = self.policy(net_in)
        mean = torch.clamp(mean.double(), -1, 1)  # ppo_agent.py:141
        std = torch.exp(self.log_std.detach().double()).expand_as(mean)
        if self.use_mask_sampling:
            action, _ = masked_discrete_actions(mean, std, obs["action_mask"], self.actions42, self.gen)
        else:
            action = torch.clamp(mean + std * torch.randn(mean.shape, dtype=mean.dtype, device=mean.device, generator=self.gen), -1, 1)
        return action.contiguous(), (mean, std)

    @staticmethod
    def log_prob(action, dist):
        """Gaussian log-density of `action` under the policy's (mean, std) (ppo_agent.py get_log_prob)"""
        mean, std = dist
        return -0.5 * ((action - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)

    @torch.no_grad()
    def collect(self, n_steps, store=None):
        """Run n_steps env steps.  `store`, if given, is called each step with (t, obs, action, reward, done, log_prob, executing):
        `obs` is the observation the policy ACTED ON (the env writes its outputs in place, so it is copied before the step when a
        store is given), `action` the action the env executed — the Reeds-Shepp plan's where a route is being executed — and
        `log_prob` the policy's log-density of that executed action (parking_agent.py:93-97 recomputes it for plan actions)."""
        env = self.env
        for t in range(n_steps):
            action, dist = self.act(self.obs)
            executing = None
            if self.use_planner:
                action, executing = env.planner_actions(action)
            acted_on = None
            if store is not None:
                acted_on = {k: (v.clone() if v is not None else None) for k, v in self.obs.items()}
                log_prob = self.log_prob(action, dist)
            obs, reward, done, info = env.step(action)
            if store is not None:
                store(t, acted_on, action, reward, done, log_prob, executing)
            self.obs = obs
        return self.obs
