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

from .rollout import FusedPolicy, ReferenceShapedActor


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
        self.capacity, self.device = int(capacity), device
        self._head = self._size = self._ar = self._trash = None
        self.pushed_upper_bound = 0  # rows offered so far (host counter; >= the number kept)
        if keys is not None:
            self.KEYS = tuple(keys)
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        rows = self.capacity + 1     # + the spare slot rows that are not kept are written to
        self._obs = {k: f(rows, d) for k, d in self.KEYS}
        self._nxt = {k: f(rows, d) for k, d in self.KEYS}
        self._action, self._reward, self._done = f(rows, 2), f(rows), f(rows)
        # the ring itself (views without the spare slot)
        self.obs = {k: v[:self.capacity] for k, v in self._obs.items()}
        self.nxt = {k: v[:self.capacity] for k, v in self._nxt.items()}
        self.action, self.reward, self.done = self._action[:self.capacity], self._reward[:self.capacity], self._done[:self.capacity]

    def push(self, obs, action, reward, done, nxt, keep=None):
        """Append the rows with keep != 0 (all rows when keep is None).  No host synchronisation: the write position and the
        fill level live on the device, rows that are not kept are written to a spare slot behind the ring."""
        n = action.shape[0]
        if self._head is None:
            self._head = torch.zeros((), dtype=torch.long, device=self.device)
            self._size = torch.zeros((), dtype=torch.long, device=self.device)
            self._ar = torch.arange(n, device=self.device)
            self._trash = torch.full((), self.capacity, dtype=torch.long, device=self.device)
        if self._ar.numel() != n:
            self._ar = torch.arange(n, device=self.device)
        if keep is not None:  # e.g. drop the auto-reset pseudo-steps
            k = keep.to(torch.bool)
            rank = torch.cumsum(k.to(torch.long), 0) - 1
            idx = torch.where(k, (rank + self._head) % self.capacity, self._trash)
            cnt = rank[-1] + 1
        else:
            idx = (self._ar + self._head) % self.capacity
            cnt = torch.full((), n, dtype=torch.long, device=self.device)
        for key, _ in self.KEYS:
            self._obs[key].index_copy_(0, idx, obs[key].float())
            self._nxt[key].index_copy_(0, idx, nxt[key].float())
        self._action.index_copy_(0, idx, action.float()); self._reward.index_copy_(0, idx, reward.float()); self._done.index_copy_(0, idx, done.float())
        self._head = (self._head + cnt) % self.capacity
        self._size = torch.clamp(self._size + cnt, max=self.capacity)
        self.pushed_upper_bound += n

    @property
    def size(self):
        """fill level (reads the device counter: synchronises; the rollout loop uses `pushed_upper_bound` instead)"""
        return 0 if self._size is None else int(self._size.item())

    def sample(self, batch, generator=None):
        size = self._size.clamp(min=1)
        idx = (torch.rand(batch, device=self.device, generator=generator) * size).long().clamp_(max=self.capacity - 1)
        o = {k: self.obs[k][idx] for k, _ in self.KEYS}
        n = {k: self.nxt[k][idx] for k, _ in self.KEYS}
        return o, self.action[idx], self.reward[idx], self.done[idx], n


class FlatGradAllReduce(object):
    """All gradients of several modules as ONE contiguous bucket -> one all-reduce per update (the message
    is a few MB, latency-bound on NVLink).  The parameters' .grad tensors ARE views into the bucket (assigned once,
    kept by zero_grad(set_to_none=False)), so there is nothing to gather or scatter: `launch` is the asynchronous all-reduce,
    `wait` averages in place."""

    def __init__(self, modules, world, extra_params=()):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad] + [p for p in extra_params if p.requires_grad]
        self.world = world
        self.numel = sum(p.numel() for p in self.params)
        self.bucket = torch.zeros(self.numel, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.bucket[off:off + n].view_as(p)
            off += n
        self.handle = None

    def launch(self):
        if self.world > 1:
            self.handle = dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self.handle is not None:
            self.handle.wait()
            self.handle = None
        if self.world > 1:
            self.bucket.div_(self.world)


class SacLite(object):
    def __init__(self, device, world=1, lr=3e-4, gamma=0.98, tau=0.005, seed=0):
        torch.manual_seed(seed)  # identical initial weights on every rank
        self.actor = ReferenceShapedActor().to(device)
        self.q1, self.q2 = QNet().to(device), QNet().to(device)
        self.q1_t, self.q2_t = QNet().to(device), QNet().to(device)
        self.q1_t.load_state_dict(self.q1.state_dict()); self.q2_t.load_state_dict(self.q2.state_dict())
        self.log_alpha = torch.zeros((), device=device, requires_grad=True)
        self.gamma, self.tau, self.target_entropy = gamma, tau, -2.0
        self.use_graph, self._graph, self._static, self.graph_error = True, None, None, None
        fused = torch.device(device).type == "cuda"   # one kernel per optimiser step instead of one per parameter group of tensors
        self.opt_actor = torch.optim.Adam(list(self.actor.parameters()), lr=lr, fused=fused)
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=lr, fused=fused)
        self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=lr, fused=fused)
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
        launches their all-reduce.  Call `apply()` later (after some rollout steps) to finish the update.
        On a CUDA device the ~800 kernels of the forward / backward passes are captured once into a CUDA graph (static batch
        buffers, gradients kept allocated) and replayed: what an update costs in eager mode is the host time to launch them."""
        if self.use_graph and batch[1].is_cuda:
            return self._backward_graphed(batch)
        out = self._forward_backward(batch)
        self.reducer.launch()
        return out

    def _backward_graphed(self, batch):
        obs, act, rew, done, nxt = batch
        if self._graph is None:
            self._static = ({k: v.clone() for k, v in obs.items()}, act.clone(), rew.clone(), done.clone(), {k: v.clone() for k, v in nxt.items()})
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=act.device)
            side.wait_stream(cur)
            try:
                with torch.cuda.stream(side):     # warm-up outside the capture (cuBLAS workspaces, autograd buffers, .grad tensors)
                    for _ in range(3):
                        self._forward_backward(self._static)
                cur.wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._forward_backward(self._static)
                self._graph = g
            except Exception as e:                # no capture on this build: eager updates from now on
                self.use_graph, self._graph, self.graph_error = False, None, repr(e)
                cur.wait_stream(side)
                torch.cuda.synchronize()
                out = self._forward_backward(batch)
                self.reducer.launch()
                return out
        so, sa, sr, sd, sn = self._static
        for k in so:
            so[k].copy_(obs[k]); sn[k].copy_(nxt[k])
        sa.copy_(act); sr.copy_(rew); sd.copy_(done)
        self._graph.replay()
        self.reducer.launch()
        return None, None

    def _forward_backward(self, batch):
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
        self.opt_alpha.zero_grad(set_to_none=False)
        (-(self.log_alpha * (logp.detach().float().mean() + self.target_entropy))).backward()
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
        with torch.no_grad():  # Polyak averaging of both target critics as one multi-tensor operation
            torch._foreach_lerp_(list(self.q1_t.parameters()) + list(self.q2_t.parameters()), list(self.q1.parameters()) + list(self.q2.parameters()), self.tau)


class SacRollout(object):
    """Rollout + learning loop of BASELINE cfg 5 on one rank: act with the tanh-Gaussian actor (RS plan
    hand-off on), push every real transition to the device replay, and every `update_every` env steps run one
    SAC update whose gradient all-reduce overlaps the following rollout steps."""

    def __init__(self, env, world=1, batch=8192, update_every=8, replay_capacity=1 << 20, seed=0, policy_kernel=True):
        self.env, self.batch, self.update_every = env, batch, update_every
        self.learner = SacLite(env.device, world, seed=seed)
        self.replay = DeviceReplay(replay_capacity, env.device)
        self.gen = torch.Generator(device=env.device); self.gen.manual_seed(1000 + seed)
        self.obs = self._f32(env.reset())
        env.planner_reset()
        self.pending_update = False
        self.updates = 0
        self.overlap, self._pol_stream, self._next_action, self._steps_done = True, None, None, 0
        # acting: the actor's forward as one kernel (csrc/policy_forward.cu), its packed weights re-read after every update
        self.actor_kernel = FusedPolicy(self.learner.actor, env.n, env.device) if (policy_kernel and FusedPolicy.supports(self.learner.actor) == 3) else None

    @torch.no_grad()
    def _act(self, obs):
        L = self.learner
        if self.actor_kernel is None:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return L._pi(obs)[0]
        mean = self.actor_kernel(obs)
        return torch.tanh(mean + torch.exp(L.actor.log_std) * torch.randn(mean.shape, dtype=mean.dtype, device=mean.device, generator=self.gen))

    @staticmethod
    def _f32(o):
        return {"lidar": o["lidar"].float(), "target": o["target"].float(), "action_mask": o["action_mask"].float()}

    def run(self, n_steps):
        """n_steps rollout steps.  The policy side of the next step (float32 casts + actor forward: it needs the observation
        only) runs on a second stream next to the Reeds-Shepp kernels of the current step (hope_wait_observed), except on the
        steps that finish an update, where the actor's new weights have to be in place first."""
        env, L = self.env, self.learner
        dev = env.device
        main = torch.cuda.current_stream(dev)
        if self._pol_stream is None:
            self._pol_stream = torch.cuda.Stream(device=dev)
            self._upd_stream = torch.cuda.Stream(device=dev)
            self._weights_ready = torch.cuda.Event()
        pol = self._pol_stream
        a = self._act(self.obs) if self._next_action is None else self._next_action
        for t in range(n_steps):
            act, executing = env.planner_actions(a.double().contiguous())
            obs, reward, done, info = env.step(act)
            update_due = (self._steps_done + 1) % self.update_every == 0 and self.replay.pushed_upper_bound + env.n >= 2 * self.batch
            ahead = self.overlap and not update_due and env.wait_observed(pol)
            if ahead:
                with torch.cuda.stream(pol):
                    nxt = self._f32(obs)
                    a = self._act(nxt)
                    for x in (a,) + tuple(nxt.values()):
                        x.record_stream(main)
                main.wait_stream(pol)
            else:
                nxt = self._f32(obs)
            self.replay.push(self.obs, act, reward, done, nxt, keep=info["was_reset"] == 0)
            self.obs = nxt
            self._steps_done += 1
            if update_due:
                # The learner has a stream of its own: the batch is gathered here (a private copy of the sampled rows, so the
                # following pushes cannot tear it), the previous update is finished (its all-reduce ran under the last rollout
                # steps) and the actor's packed weights are refreshed; the rollout only waits for THAT, while the forward /
                # backward passes of the new update and their all-reduce run under the next update_every rollout steps.
                batch = self.replay.sample(self.batch, self.gen)
                upd = self._upd_stream
                upd.wait_stream(main)
                with torch.cuda.stream(upd):
                    for x in list(batch[0].values()) + list(batch[4].values()) + [batch[1], batch[2], batch[3]]:
                        x.record_stream(upd)
                    if self.pending_update:
                        L.apply()
                        if self.actor_kernel is not None:
                            self.actor_kernel.refresh()
                    self._weights_ready.record(upd)
                    L.backward(batch)
                main.wait_event(self._weights_ready)
                self.pending_update = True
                self.updates += 1
            if not ahead:
                a = self._act(self.obs)
        self._next_action = a
        if self.pending_update:
            with torch.cuda.stream(self._upd_stream):
                L.apply()
                if self.actor_kernel is not None:
                    self.actor_kernel.refresh()
            main.wait_stream(self._upd_stream)
            self.pending_update = False
            self._next_action = None   # computed with the weights before this update
