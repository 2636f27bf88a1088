"""Minimal SAC-style learner around the device-resident rollout (BASELINE cfg 5): a device replay
buffer, twin critics, learned temperature, and ONE flat NCCL all-reduce of all gradients per update.

The reference's agents (src/model/agent/sac_agent.py) are out of scope and stay untouched; this module
exists so the multi-GPU configuration can be measured end to end: scenes shard by rank with no
data-path collective, the only exchange is the gradient all-reduce (sum / world) over NVLink, launched
asynchronously so the next rollout steps overlap it.  Network shapes follow ACTOR_CONFIGS /
CRITIC_CONFIGS (configs.py:134-176) with the image modality off.
"""
import math

import torch
from torch import nn
import torch.distributed as dist

from .rollout import ReferenceShapedActor


class QNet(nn.Module):
    """MultiObsEmbedding-shaped critic: lidar / target / mask / action tokens -> scalar (sac_agent.py critics)."""

    def __init__(self, embed=128, heads=8, dim_head=32, mlp=128, hidden=128):
        super().__init__()
        emb = lambda d: nn.Sequential(nn.Linear(d, embed), nn.Tanh(), nn.Linear(embed, embed))
        self.e_lidar, self.e_tgt, self.e_am, self.e_act = emb(120), emb(5), emb(42), emb(2)
        self.norm1, self.norm2 = nn.LayerNorm(embed), nn.LayerNorm(embed)
        self.heads, self.dim_head = heads, dim_head
        self.to_qkv = nn.Linear(embed, heads * dim_head * 3, bias=False)
        self.to_out = nn.Linear(heads * dim_head, embed)
        self.ff = nn.Sequential(nn.Linear(embed, mlp), nn.Tanh(), nn.Linear(mlp, embed))
        self.head = nn.Sequential(nn.Linear(4 * embed, hidden), nn.Tanh(), nn.Linear(hidden, 1))

    def forward(self, obs, action):
        x = torch.stack([self.e_lidar(obs["lidar"]), self.e_tgt(obs["target"]), self.e_am(obs["action_mask"]), self.e_act(action)], dim=1)
        b, n, _ = x.shape
        q, k, v = self.to_qkv(self.norm1(x)).view(b, n, 3, self.heads, self.dim_head).permute(2, 0, 3, 1, 4)
        a = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (self.dim_head ** -0.5), dim=-1), v)
        x = self.to_out(a.transpose(1, 2).reshape(b, n, self.heads * self.dim_head)) + x
        x = self.ff(self.norm2(x)) + x
        return self.head(x.reshape(b, -1)).squeeze(-1)


class DeviceReplay(object):
    """Ring buffer of transitions in HBM (replay_memory.py:6-50 batched): float32 observations as the
    networks consume them, pushed N at a time straight from the env's output tensors."""

    KEYS = (("lidar", 120), ("target", 5), ("action_mask", 42))

    def __init__(self, capacity, device, keys=None):
        self.capacity, self.device, self.size, self.head = int(capacity), device, 0, 0
        if keys is not None:
            self.KEYS = tuple(keys)
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.obs = {k: f(capacity, d) for k, d in self.KEYS}
        self.nxt = {k: f(capacity, d) for k, d in self.KEYS}
        self.action, self.reward, self.done = f(capacity, 2), f(capacity), f(capacity)

    def push(self, obs, action, reward, done, nxt, keep=None):
        n = action.shape[0]
        idx = (torch.arange(n, device=self.device) + self.head) % self.capacity
        if keep is not None:  # e.g. drop the auto-reset pseudo-steps
            sel = keep.nonzero(as_tuple=True)[0]
            idx = (torch.arange(sel.numel(), device=self.device) + self.head) % self.capacity
            n = sel.numel()
            pick = lambda t: t[sel]
        else:
            pick = lambda t: t
        for k, _ in self.KEYS:
            self.obs[k][idx] = pick(obs[k]).float()
            self.nxt[k][idx] = pick(nxt[k]).float()
        self.action[idx] = pick(action).float(); self.reward[idx] = pick(reward).float(); self.done[idx] = pick(done).float()
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch, generator=None):
        idx = torch.randint(0, max(self.size, 1), (batch,), device=self.device, generator=generator)
        o = {k: self.obs[k][idx] for k, _ in self.KEYS}
        n = {k: self.nxt[k][idx] for k, _ in self.KEYS}
        return o, self.action[idx], self.reward[idx], self.done[idx], n


class FlatGradAllReduce(object):
    """All gradients of several modules as ONE contiguous bucket -> one all-reduce per update (the message
    is a few MB, latency-bound on NVLink).  `launch` is asynchronous; `wait` averages and
    scatters the result back into the .grad tensors."""

    def __init__(self, modules, world, extra_params=()):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad] + [p for p in extra_params if p.requires_grad]
        self.world = world
        self.numel = sum(p.numel() for p in self.params)
        self.bucket = torch.zeros(self.numel, dtype=torch.float32, device=self.params[0].device)
        self.handle = None

    def launch(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.bucket[off:off + n].zero_()
            else:
                self.bucket[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if self.world > 1:
            self.handle = dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self.handle is not None:
            self.handle.wait()
            self.handle = None
        if self.world > 1:
            self.bucket.div_(self.world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                p.grad.copy_(self.bucket[off:off + n].view_as(p.grad))
            off += n


class SacLite(object):
    def __init__(self, device, world=1, lr=3e-4, gamma=0.98, tau=0.005, seed=0):
        torch.manual_seed(seed)  # identical initial weights on every rank
        self.actor = ReferenceShapedActor().to(device)
        self.q1, self.q2 = QNet().to(device), QNet().to(device)
        self.q1_t, self.q2_t = QNet().to(device), QNet().to(device)
        self.q1_t.load_state_dict(self.q1.state_dict()); self.q2_t.load_state_dict(self.q2.state_dict())
        self.log_alpha = torch.zeros((), device=device, requires_grad=True)
        self.gamma, self.tau, self.target_entropy = gamma, tau, -2.0
        self.opt_actor = torch.optim.Adam(list(self.actor.parameters()), lr=lr)
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=lr)
        self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=lr)
        # the temperature is a replicated parameter too: its gradient rides in the same bucket, or the ranks' alphas drift apart
        self.reducer = FlatGradAllReduce([self.actor, self.q1, self.q2], world, extra_params=[self.log_alpha])

    def _pi(self, obs):
        mean = self.actor(obs)
        std = torch.exp(self.actor.log_std).expand_as(mean)
        u = mean + std * torch.randn_like(mean)
        a = torch.tanh(u)
        logp = (-0.5 * ((u - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)).sum(-1) - torch.log(1 - a * a + 1e-6).sum(-1)
        return a, logp

    def backward(self, batch):
        """Forward + backward of critic, actor and temperature losses; leaves gradients in .grad and
        launches their all-reduce.  Call `apply()` later (after some rollout steps) to finish the update."""
        obs, act, rew, done, nxt = batch
        alpha = self.log_alpha.exp().detach()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=obs["lidar"].is_cuda):
            with torch.no_grad():
                na, nlogp = self._pi(nxt)
                tq = torch.min(self.q1_t(nxt, na), self.q2_t(nxt, na)).float() - alpha * nlogp.float()
                y = rew + self.gamma * (1 - done) * tq
            lq = ((self.q1(obs, act).float() - y) ** 2).mean() + ((self.q2(obs, act).float() - y) ** 2).mean()
        self.opt_q.zero_grad(set_to_none=False); self.opt_actor.zero_grad(set_to_none=False)
        lq.backward()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=obs["lidar"].is_cuda):
            a, logp = self._pi(obs)
            for p in list(self.q1.parameters()) + list(self.q2.parameters()):
                p.requires_grad_(False)
            la = (alpha * logp.float() - torch.min(self.q1(obs, a), self.q2(obs, a)).float()).mean()
        la.backward()
        for p in list(self.q1.parameters()) + list(self.q2.parameters()):
            p.requires_grad_(True)
        self.opt_alpha.zero_grad()
        (-(self.log_alpha * (logp.detach().float().mean() + self.target_entropy))).backward()
        self.reducer.launch()
        return lq.detach(), la.detach()

    def time_allreduce(self, reps=5):
        """milliseconds of the flat gradient all-reduce alone (launch + wait, nothing overlapping), CUDA events"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.reducer.launch(); self.reducer.wait()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            self.reducer.launch(); self.reducer.wait()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def apply(self):
        self.reducer.wait()
        self.opt_q.step(); self.opt_actor.step(); self.opt_alpha.step()
        with torch.no_grad():
            for t, s in ((self.q1_t, self.q1), (self.q2_t, self.q2)):
                for pt, ps in zip(t.parameters(), s.parameters()):
                    pt.lerp_(ps, self.tau)


class SacRollout(object):
    """Rollout + learning loop of BASELINE cfg 5 on one rank: act with the tanh-Gaussian actor (RS plan
    hand-off on), push every real transition to the device replay, and every `update_every` env steps run one
    SAC update whose gradient all-reduce overlaps the following rollout steps."""

    def __init__(self, env, world=1, batch=8192, update_every=8, replay_capacity=1 << 20, seed=0):
        self.env, self.batch, self.update_every = env, batch, update_every
        self.learner = SacLite(env.device, world, seed=seed)
        self.replay = DeviceReplay(replay_capacity, env.device)
        self.gen = torch.Generator(device=env.device); self.gen.manual_seed(1000 + seed)
        self.obs = self._f32(env.reset())
        env.planner_reset()
        self.pending_update = False
        self.updates = 0

    @staticmethod
    def _f32(o):
        return {"lidar": o["lidar"].float(), "target": o["target"].float(), "action_mask": o["action_mask"].float()}

    def run(self, n_steps):
        env, L = self.env, self.learner
        for t in range(n_steps):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                a, _ = L._pi(self.obs)
            act, executing = env.planner_actions(a.double().contiguous())
            obs, reward, done, info = env.step(act)
            nxt = self._f32(obs)
            self.replay.push(self.obs, act, reward, done, nxt, keep=info["was_reset"] == 0)
            self.obs = nxt
            if (t + 1) % self.update_every == 0 and self.replay.size >= self.batch:
                if self.pending_update:
                    L.apply()            # finishes the previous update: its all-reduce ran under the last rollout steps
                L.backward(self.replay.sample(self.batch, self.gen))
                self.pending_update = True
                self.updates += 1
        if self.pending_update:
            L.apply()
            self.pending_update = False
