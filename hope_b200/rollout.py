"""GPU-resident rollout loop around BatchedParkingEnv — the batched counterpart of the reference's
episode loop (src/train/train_HOPE_ppo.py:190-219, train_HOPE_sac.py:191-225, scope row a19) with the
policy-side helpers that are batch-1 host code in the reference (row f2):

  masked_discrete_actions   ActionMask.choose_action     model/action_mask.py:199-227
  env.planner_actions       RsPlanner / ParkingAgent     model/agent/parking_agent.py:2-47, 60-110
  RunningNorm               StateNorm (Welford)          model/state_norm.py:25-46

The policy itself is the reference's (`model.network.MultiObsEmbedding` works with batch > 1); for
synthetic benchmarks without the reference tree, `ReferenceShapedActor` has the same layer shapes
(ACTOR_CONFIGS, configs.py:134-152, three modalities when the image is off).  The 3-modal actor's forward is one kernel
(FusedPolicy -> hope_policy_forward, bf16 tensor-core GEMMs); other policies run through stock PyTorch (bf16 autocast);
everything else stays in float64 device tensors.
"""
import math

import ctypes as C

import torch
from torch import nn

from . import capi, tables

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


class FusedStateNorm(object):
    """RunningNorm as one C-ABI call (hope_state_norm, csrc/policy_glue.cu): batch Welford + Chan merge of the running
    statistics, normalisation and the float32 cast for the network in three small launches instead of ~20 elementwise
    PyTorch kernels over (N, 120) float64 tensors.  Same statistics as RunningNorm (tests/test_gpu_rollout.py)."""

    def __init__(self, n_envs, device):
        self.lib = capi.load_library()
        self.device, self.n_envs, self.n = device, int(n_envs), 0
        self.stats = torch.zeros((2, 125), dtype=torch.float64, device=device)   # mean, M2 of [lidar | target]
        self.scratch = torch.zeros(self.lib.hope_state_norm_scratch_bytes(self.n_envs), dtype=torch.uint8, device=device)
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.out = {"lidar": f32(self.n_envs, 120), "target": f32(self.n_envs, 5), "action_mask": f32(self.n_envs, 42)}

    @property
    def mean(self):
        return {"lidar": self.stats[0, :120], "target": self.stats[0, 120:]}

    @property
    def m2(self):
        return {"lidar": self.stats[1, :120], "target": self.stats[1, 120:]}

    def __call__(self, obs, update=True):
        """obs: the env's float64 lidar / target / action_mask tensors -> float32 network inputs (buffers reused every call)"""
        n = obs["lidar"].shape[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.hope_state_norm(obs["lidar"].data_ptr(), obs["target"].data_ptr(), obs["action_mask"].data_ptr(), n,
                                            self.stats.data_ptr(), float(self.n), 1 if update else 0, self.scratch.data_ptr(),
                                            self.out["lidar"].data_ptr(), self.out["target"].data_ptr(), self.out["action_mask"].data_ptr(), stream))
        if update:
            self.n += n
        return self.out


class FusedMaskedSampler(object):
    """masked_discrete_actions as one kernel (hope_masked_sample): 42 clipped Gaussian log-densities x mask per env and an
    inverse-CDF draw from a Philox stream keyed by (seed, step, env)."""

    def __init__(self, n_envs, device, seed=0):
        self.lib = capi.load_library()
        self.device, self.seed, self.step = device, int(seed), 0
        self.actions42 = possible_actions(device).contiguous()
        self.action = torch.zeros((n_envs, 2), dtype=torch.float64, device=device)
        self.index = torch.zeros(n_envs, dtype=torch.int32, device=device)
        self.u = torch.zeros(n_envs, dtype=torch.float64, device=device)

    def __call__(self, mean_f32, log_std, mask):
        n = mean_f32.shape[0]
        assert mean_f32.dtype == torch.float32 and mean_f32.is_contiguous() and mask.dtype == torch.float64
        ls = log_std.detach().double().contiguous()
        capi.check(self.lib.hope_masked_sample(n, mean_f32.data_ptr(), ls.data_ptr(), mask.data_ptr(), self.actions42.data_ptr(), self.seed, self.step,
                                               self.action.data_ptr(), self.index.data_ptr(), self.u.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        self.step += 1
        return self.action, self.index


class FusedPolicy(object):
    """The actor (MultiObsEmbedding(ACTOR_CONFIGS) / ReferenceShapedActor: lidar, target, action mask, and with the 4-modal
    network the image token from its encoder's mean head) as one kernel, hope_policy_forward[_img] (csrc/policy_forward.cu):
    32 (4-modal: 16) envs per CTA carried through the embeddings, the transformer block and the output head with every
    activation in shared memory, bf16 tensor-core GEMMs with float32 accumulation.  The parameters are
    read from the module's state_dict by the reference's names and packed once (`refresh()` after an optimiser step): bf16, input
    width padded to 16, in the kernel's fragment order (include/hope_b200.h, hope_policy_pack_matrix)."""

    _MATS = (("w1_lidar", "embed_lidar.0.weight", 128), ("w1_target", "embed_tgt.0.weight", 16), ("w1_mask", "embed_am.0.weight", 48),
             ("w2_0", "embed_lidar.2.weight", 128), ("w2_1", "embed_tgt.2.weight", 128), ("w2_2", "embed_am.2.weight", 128),
             ("w_qkv", "net.encoder.layers.0.0.fn.to_qkv.weight", 128), ("w_out", "net.encoder.layers.0.0.fn.to_out.0.weight", 256),
             ("w_ff1", "net.encoder.layers.0.1.fn.net.0.weight", 128), ("w_ff2", "net.encoder.layers.0.1.fn.net.3.weight", 128),
             ("w_o1", "net.output.0.weight", 384))
    _VECS = (("b1_0", "embed_lidar.0.bias"), ("b1_1", "embed_tgt.0.bias"), ("b1_2", "embed_am.0.bias"),
             ("b2_0", "embed_lidar.2.bias"), ("b2_1", "embed_tgt.2.bias"), ("b2_2", "embed_am.2.bias"),
             ("ln1_g", "net.encoder.layers.0.0.norm.weight"), ("ln1_b", "net.encoder.layers.0.0.norm.bias"),
             ("b_out", "net.encoder.layers.0.0.fn.to_out.0.bias"),
             ("ln2_g", "net.encoder.layers.0.1.norm.weight"), ("ln2_b", "net.encoder.layers.0.1.norm.bias"),
             ("b_ff1", "net.encoder.layers.0.1.fn.net.0.bias"), ("b_ff2", "net.encoder.layers.0.1.fn.net.3.bias"),
             ("b_o1", "net.output.0.bias"), ("w_o2", "net.output.2.weight"), ("b_o2", "net.output.2.bias"))
    _SHAPES = {"embed_lidar.0.weight": (128, 120), "embed_tgt.0.weight": (128, 5), "embed_am.0.weight": (128, 42),
               "net.encoder.layers.0.0.fn.to_qkv.weight": (768, 128), "net.encoder.layers.0.0.fn.to_out.0.weight": (128, 256),
               "net.output.0.weight": (128, 384), "net.output.2.weight": (2, 128)}

    _IMG_MATS = (("w2_img", "re_embed_img.1.weight", 128),)
    _IMG_VECS = (("b2_img", "re_embed_img.1.bias"),)

    @classmethod
    def supports(cls, module):
        """3 when `module` has exactly the lidar + target + action-mask architecture the kernel implements, 4 when it is the
        4-modal network (the same plus the image token: embed_img / re_embed_img, Linear(512, 128) head) and exposes the image
        encoder's tail (`img_mean_from_conv`), else 0"""
        sd = module.state_dict()
        names = [m[1] for m in cls._MATS] + [v[1] for v in cls._VECS]
        if not all(k in sd for k in names):
            return 0
        has_img = any(k.startswith(("embed_img", "re_embed_img")) for k in sd)
        shapes = dict(cls._SHAPES)
        if has_img:
            shapes["net.output.0.weight"] = (128, 512)
            shapes["re_embed_img.1.weight"] = (128, 128)
            if "re_embed_img.1.weight" not in sd or not hasattr(module, "img_mean_from_conv"):
                return 0
        if not all(tuple(sd[k].shape) == shp for k, shp in shapes.items()):
            return 0
        return 4 if has_img else 3

    def __init__(self, module, n_envs, device):
        self.n_modal = self.supports(module)
        assert self.n_modal, "hope_policy_forward implements the 3- and 4-modal actors of ACTOR_CONFIGS only"
        self.lib = capi.load_library()
        self.module, self.device = module, device
        self.out = torch.zeros((n_envs, 2), dtype=torch.float32, device=device)
        self.weights = capi.PolicyWeights()
        self._keep = {}
        self.mats = tuple((n, k, 512 if (n == "w_o1" and self.n_modal == 4) else kp) for n, k, kp in self._MATS) + (self._IMG_MATS if self.n_modal == 4 else ())
        self.vecs = self._VECS + (self._IMG_VECS if self.n_modal == 4 else ())
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """copy the module's parameters into the kernel's packed buffers, in place (the buffers are allocated once: a forward
        that is still in flight on another stream never sees freed memory; the caller orders refresh against forwards)"""
        sd = self.module.state_dict()
        keep = self._keep
        for name, key, kpad in self.mats:
            w = sd[key].detach()
            if name not in keep:
                keep[name] = torch.zeros((w.shape[0], kpad), dtype=torch.bfloat16, device=self.device)
                keep["_pad_" + name] = torch.zeros((w.shape[0], kpad), dtype=torch.bfloat16, device=self.device)
            pad = keep["_pad_" + name]
            pad[:, :w.shape[1]].copy_(w)          # float32 -> bf16 (round to nearest even), zero-padded input width
            # fragment packing (hope_policy_pack_matrix on the device): [nt][8 rows][ks][half][4 lane%4][2] -> [nt][ks][row][lane%4][half][2]
            n_out = w.shape[0]
            keep[name].view(n_out // 8, kpad // 16, 8, 4, 2, 2).copy_(pad.view(n_out // 8, 8, kpad // 16, 2, 4, 2).permute(0, 2, 1, 4, 3, 5))
        for name, key in self.vecs:
            v = sd[key].detach()
            if name not in keep:
                keep[name] = torch.zeros(tuple(v.shape), dtype=torch.float32, device=self.device)
            keep[name].copy_(v)
        W = self.weights
        W.w1_lidar, W.w1_target, W.w1_mask = (keep[k].data_ptr() for k in ("w1_lidar", "w1_target", "w1_mask"))
        for m in range(3):
            W.w2[m] = keep[f"w2_{m}"].data_ptr(); W.b1[m] = keep[f"b1_{m}"].data_ptr(); W.b2[m] = keep[f"b2_{m}"].data_ptr()
        for k in ("w_qkv", "w_out", "w_ff1", "w_ff2", "w_o1", "ln1_g", "ln1_b", "b_out", "ln2_g", "ln2_b", "b_ff1", "b_ff2", "b_o1", "w_o2", "b_o2"):
            setattr(W, k, keep[k].data_ptr())
        if self.n_modal == 4:
            W.w2_img, W.b2_img = keep["w2_img"].data_ptr(), keep["b2_img"].data_ptr()

    def __call__(self, net_in):
        """net_in: float32 lidar (N,120), target (N,5), action_mask (N,42) -> float32 (N,2) policy mean in [-1,1] (buffer reused)"""
        lidar, target, mask = net_in["lidar"], net_in["target"], net_in["action_mask"]
        n = lidar.shape[0]
        assert n <= self.out.shape[0] and all(t.dtype == torch.float32 and t.is_contiguous() for t in (lidar, target, mask))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.n_modal == 4:  # net_in["img_mean"]: float32 (N, 128), the image encoder's mean head (module.img_mean_from_conv)
            mean = net_in["img_mean"]
            assert mean.dtype == torch.float32 and mean.is_contiguous() and tuple(mean.shape) == (n, 128)
            capi.check(self.lib.hope_policy_forward_img(n, lidar.data_ptr(), target.data_ptr(), mask.data_ptr(), mean.data_ptr(), C.byref(self.weights),
                                                        self.out.data_ptr(), stream))
        else:
            capi.check(self.lib.hope_policy_forward(n, lidar.data_ptr(), target.data_ptr(), mask.data_ptr(), C.byref(self.weights), self.out.data_ptr(), stream))
        return self.out[:n]


class FusedImgConv(object):
    """The two residual conv blocks of the 4-modal actor's image encoder (embed_img.net[0:3]: conv blocks + flatten) as one
    kernel, hope_img_conv_forward (csrc/img_encoder.cu): one CTA per image, both blocks in shared memory, uint8 image in,
    bf16 features (N, 2048) out.  The 464 weights travel by value in the kernel's parameter block; `refresh()` re-reads them."""

    _KEYS = (("conv1_w", "embed_img.net.0.layer.0.weight", (4, 3, 3, 3)), ("conv1_b", "embed_img.net.0.layer.0.bias", (4,)),
             ("short1_w", "embed_img.net.0.shortcut.0.weight", (4, 3, 1, 1)), ("short1_b", "embed_img.net.0.shortcut.0.bias", (4,)),
             ("conv2_w", "embed_img.net.1.layer.0.weight", (8, 4, 3, 3)), ("conv2_b", "embed_img.net.1.layer.0.bias", (8,)),
             ("short2_w", "embed_img.net.1.shortcut.0.weight", (8, 4, 1, 1)), ("short2_b", "embed_img.net.1.shortcut.0.bias", (8,)))

    @classmethod
    def supports(cls, module):
        sd = module.state_dict()
        return all(k in sd and tuple(sd[k].shape) == shp for _, k, shp in cls._KEYS) and hasattr(module, "forward_from_img_features")

    def __init__(self, module, n_envs, device):
        assert self.supports(module)
        self.lib = capi.load_library()
        self.module, self.device = module, device
        self.feat = torch.zeros((n_envs, 2048), dtype=torch.bfloat16, device=device)
        self.weights = capi.ImgConvWeights()
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        sd = self.module.state_dict()
        flat = torch.cat([sd[k].detach().float().reshape(-1) for _, k, _ in self._KEYS]).cpu().numpy()   # 464 floats, one copy
        C.memmove(C.addressof(self.weights), flat.ctypes.data, flat.nbytes)

    def __call__(self, img_u8):
        n = img_u8.shape[0]
        assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and tuple(img_u8.shape[1:]) == (3, 64, 64) and n <= self.feat.shape[0]
        capi.check(self.lib.hope_img_conv_forward(n, img_u8.data_ptr(), C.byref(self.weights), self.feat.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return self.feat if n == self.feat.shape[0] else self.feat[:n]


class _ConvBlock(nn.Module):
    """network.py:198-232 with the shipped switches (no batch norm, residual on, tanh): conv3x3 -> tanh -> maxpool2, plus the
    conv1x1 -> avgpool2 shortcut"""

    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.layer = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2), nn.Tanh(), nn.MaxPool2d(2))
        self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1), nn.AvgPool2d(2))

    def forward(self, x):
        return self.layer(x) + self.shortcut(x)


class _ImgEncoder(nn.Module):
    """network.py:278-299: two conv blocks (3 -> 4 -> 8 channels), flatten, Linear(2048, 256), tanh, mean / std heads"""

    def __init__(self, shape=(3, 64, 64), convs=(4, 8), fc=256, embed=128):
        super().__init__()
        c, w, h = shape
        layers, cin = [], c
        for cout in convs:
            layers.append(_ConvBlock(cin, cout))
            cin = cout
        layers += [nn.Flatten(), nn.Linear(w * h * convs[-1] // (4 ** len(convs)), fc), nn.Tanh()]
        self.net = nn.Sequential(*layers)
        self.output_mean, self.output_std = nn.Linear(fc, embed), nn.Linear(fc, embed)

    def forward(self, x):
        x = self.net(x)
        return self.output_mean(x), self.output_std(x)


class ReferenceShapedActor(nn.Module):
    """The reference's actor, `MultiObsEmbedding(ACTOR_CONFIGS)` (network.py:34-196, attention.py:16-92, configs.py:134-153),
    restated so that it runs where the reference tree is not importable (the GPU box): per-modality 2-layer tanh embeddings
    (lidar, target, action mask and — with use_img — the ImgEncoder conv stack + re-embedding) -> n_modal x 128 tokens -> one
    pre-norm transformer block (8 heads x 32, FF 128) -> Linear(n_modal * 128, 128) tanh Linear(128, 2) tanh.
    Parameter names follow the reference module tree, so `load_state_dict(MultiObsEmbedding(...).state_dict())` and the
    `actor_net` of the shipped checkpoints load as they are (tests/test_policy_shape.py checks outputs against the real class)."""

    def __init__(self, lidar=120, target=5, mask=42, embed=128, heads=8, dim_head=32, mlp=128, hidden=128, out=2, use_img=False):
        super().__init__()
        emb = lambda d: nn.Sequential(nn.Linear(d, embed), nn.Tanh(), nn.Linear(embed, embed))
        self.embed_lidar, self.embed_tgt, self.embed_am = emb(lidar), emb(target), emb(mask)
        self.use_img = bool(use_img)
        if self.use_img:
            self.embed_img = _ImgEncoder(embed=embed)
            self.re_embed_img = nn.Sequential(nn.Tanh(), nn.Linear(embed, embed))
        self.n_modal = 3 + int(self.use_img)
        self.heads, self.dim_head = heads, dim_head
        inner = heads * dim_head
        attn = nn.Module()
        attn.norm = nn.LayerNorm(embed)
        attn.fn = nn.Module()
        attn.fn.to_qkv = nn.Linear(embed, inner * 3, bias=False)
        attn.fn.to_out = nn.Sequential(nn.Linear(inner, embed), nn.Identity())
        ff = nn.Module()
        ff.norm = nn.LayerNorm(embed)
        ff.fn = nn.Module()
        ff.fn.net = nn.Sequential(nn.Linear(embed, mlp), nn.Tanh(), nn.Identity(), nn.Linear(mlp, embed), nn.Identity())
        self.net = nn.Module()
        self.net.encoder = nn.Module()
        self.net.encoder.layers = nn.ModuleList([nn.ModuleList([attn, ff])])
        self.net.output = nn.Sequential(nn.Linear(self.n_modal * embed, hidden), nn.Tanh(), nn.Linear(hidden, out))
        self.net.view_embed = nn.Parameter(torch.zeros(1, self.n_modal, embed))  # present (and unused) in the reference too
        self.log_std = nn.Parameter(torch.zeros(out))  # the agents keep it next to the net (ppo_agent.py / sac_agent.py); not a key of the net

    def forward(self, obs):
        feats = [self.embed_lidar(obs["lidar"]), self.embed_tgt(obs["target"]), self.embed_am(obs["action_mask"])]
        if self.use_img:
            feats.append(self.re_embed_img(self.embed_img(obs["img"])[0]))
        return self._from_tokens(feats)

    def forward_from_img_features(self, obs, conv_feat):
        """forward() with the conv stack of the image encoder already applied: conv_feat (N, 2048) = embed_img.net[0:3](img)
        (FusedImgConv); the image encoder's Linear / tanh / mean head and everything after them run here"""
        feats = [self.embed_lidar(obs["lidar"]), self.embed_tgt(obs["target"]), self.embed_am(obs["action_mask"])]
        feats.append(self.re_embed_img(self.img_mean_from_conv(conv_feat)))
        return self._from_tokens(feats)

    def img_mean_from_conv(self, conv_feat):
        """the image encoder after its conv stack: Linear(2048, 256), tanh, mean head -> (N, 128) (what re_embed_img consumes)"""
        return self.embed_img.output_mean(self.embed_img.net[3:](conv_feat))

    def _from_tokens(self, feats):
        x = torch.stack(feats, dim=1)
        b, n, _ = x.shape
        attn, ff = self.net.encoder.layers[0]
        q, k, v = attn.fn.to_qkv(attn.norm(x)).view(b, n, 3, self.heads, self.dim_head).permute(2, 0, 3, 1, 4)
        # n_modal tokens per env: explicit softmax(q k^T / sqrt(d)) v like attention.py:33-46 (the fused SDPA back ends
        # reject a 65 536 x 8-head batch of 3-token sequences)
        a = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (self.dim_head ** -0.5), dim=-1), v)
        x = attn.fn.to_out(a.transpose(1, 2).reshape(b, n, self.heads * self.dim_head)) + x
        x = ff.fn.net(ff.norm(x)) + x
        return torch.tanh(self.net.output(x.reshape(b, n * x.shape[-1])))


def reference_actor(use_img=False, device=None):
    """The policy of BASELINE cfg 4: the reference's own `MultiObsEmbedding(ACTOR_CONFIGS)` when its tree is importable
    (src/ on sys.path), else the restatement above.  Returns (module, description)."""
    try:
        from model.network import MultiObsEmbedding  # the reference's src/model, untouched
        import configs
        cfg = dict(configs.ACTOR_CONFIGS)
        cfg["img_shape"] = (3, 64, 64) if use_img else None
        cfg["n_modal"] = 3 + int(use_img)
        net, what = MultiObsEmbedding(cfg), "model.network.MultiObsEmbedding(ACTOR_CONFIGS) from the reference tree"
    except Exception:
        net, what = ReferenceShapedActor(use_img=use_img), "ReferenceShapedActor (restatement of MultiObsEmbedding(ACTOR_CONFIGS); reference tree not importable here)"
    if not hasattr(net, "log_std"):
        net.log_std = nn.Parameter(torch.zeros(2))
    return (net.to(device) if device is not None else net), what


class RolloutEngine(object):
    """PPO-style acting loop for N envs: normalise -> policy (bf16 autocast) -> masked discrete sampling ->
    RS plan override -> env.step, all on the device; no host synchronisation inside `collect`."""

    def __init__(self, env, policy, log_std=None, use_planner=True, use_mask_sampling=True, state_norm=True,
                 autocast_dtype=torch.bfloat16, seed=0, fused=True, graph=True, policy_kernel=True, overlap=True):
        """fused: state norm and masked sampling through the CUDA kernels of csrc/policy_glue.cu (False: the eager PyTorch
        versions above, kept as their numerics reference).  policy_kernel: a 3-modal actor runs as the one-kernel forward of
        csrc/policy_forward.cu (call `refresh_policy()` after changing its parameters); otherwise graph: the PyTorch forward
        is captured once into a CUDA graph and replayed (one launch instead of ~60 small kernels per step)."""
        self.env, self.policy = env, policy
        dev = env.device
        self.actions42 = possible_actions(dev)
        self.use_planner, self.use_mask_sampling = use_planner, use_mask_sampling
        self.fused = bool(fused) and state_norm and use_mask_sampling
        self.norm = (FusedStateNorm(env.n, dev) if self.fused else RunningNorm({"lidar": (120,), "target": (5,)}, dev)) if state_norm else None
        self.sampler = FusedMaskedSampler(env.n, dev, seed) if self.fused else None
        self.log_std = log_std if log_std is not None else getattr(policy, "log_std", torch.zeros(2, device=dev))
        self.autocast_dtype = autocast_dtype
        self.gen = torch.Generator(device=dev); self.gen.manual_seed(seed)
        self.use_img = env.use_img and getattr(policy, "use_img", False)
        self._graph, self._graph_in, self._graph_out = None, None, None
        self.overlap, self._pol_stream = bool(overlap), None  # collect(store=None): next action computed next to the Reeds-Shepp kernels
        self.use_graph = bool(graph) and self.fused
        self.img_conv = FusedImgConv(policy, env.n, dev) if (policy_kernel and self.fused and self.use_img and FusedImgConv.supports(policy)) else None
        want = 4 if self.use_img else 3   # (a 4-modal network on an env without images is not something the kernel path handles)
        ok = policy_kernel and self.fused and FusedPolicy.supports(policy) == want and (self.img_conv is not None or not self.use_img)
        self.policy_kernel = FusedPolicy(policy, env.n, dev) if ok else None
        self.glue = ("hope_state_norm + hope_masked_sample kernels (csrc/policy_glue.cu)" if self.fused else "eager PyTorch") + \
                    (", policy forward = hope_policy_forward (csrc/policy_forward.cu, one kernel)" if self.policy_kernel is not None else
                     (", policy forward replayed from a CUDA graph" if self.use_graph else "")) + \
                    (", image conv stack = hope_img_conv_forward (csrc/img_encoder.cu)" if self.img_conv is not None else "")
        self.obs = env.reset()
        if use_planner:
            env.planner_reset()

    def _img_mean(self, conv_feat):
        """the image encoder's tail on the conv kernel's features (two small library GEMMs under autocast) -> float32 (N, 128)"""
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            return self.policy.img_mean_from_conv(conv_feat).float().contiguous()

    def _forward(self, net_in):
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            if "img_feat" in net_in:
                return self.policy.forward_from_img_features(net_in, net_in["img_feat"]).float()
            return self.policy(net_in).float()

    def refresh_policy(self):
        """re-read the policy's parameters into the kernel's packed copy (after an optimiser step or load_state_dict)"""
        if self.policy_kernel is not None:
            self.policy_kernel.refresh()
        if self.img_conv is not None:
            self.img_conv.refresh()

    def _policy_mean(self, net_in):
        """float32 policy output for the (persistent) float32 input buffers `net_in`"""
        if self.policy_kernel is not None:
            if self.policy_kernel.n_modal == 4:
                net_in["img_mean"] = self._img_mean(net_in["img_feat"])
            return self.policy_kernel(net_in)
        if not self.use_graph:
            return self._forward(net_in)
        if self._graph is None:
            self._graph_in = net_in
            side = torch.cuda.Stream(device=self.env.device)
            side.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(side):  # warm-up outside the capture (cuBLAS workspaces, autocast weight casts)
                for _ in range(2):
                    self._forward(net_in)
            torch.cuda.current_stream(self.env.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._graph_out = self._forward(net_in)
        assert all(net_in[k] is self._graph_in[k] for k in net_in), "the captured policy reads the persistent input buffers"
        self._graph.replay()
        return self._graph_out

    @torch.no_grad()
    def act(self, obs):
        if self.fused:
            net_in = dict(self.norm(obs))                      # float32 lidar / target / action_mask, statistics updated
            if self.img_conv is not None:
                net_in["img_feat"] = self.img_conv(obs["img"])       # uint8 image -> conv features, one kernel (persistent buffer)
            elif self.use_img:
                if not hasattr(self, "_img_f32"):
                    self._img_f32 = torch.zeros(obs["img"].shape, dtype=torch.float32, device=self.env.device)
                torch.mul(obs["img"], 1.0 / 255.0, out=self._img_f32)   # the reference's float image (observation_processor.py:13-17)
                net_in["img"] = self._img_f32
            mean32 = self._policy_mean(net_in).contiguous()
            action, _ = self.sampler(mean32, self.log_std, obs["action_mask"])
            mean = torch.clamp(mean32.double(), -1, 1)
            std = torch.exp(self.log_std.detach().double()).expand_as(mean)
            return action, (mean, std)
        o = {"lidar": obs["lidar"], "target": obs["target"]}
        if self.norm is not None:
            self.norm.update(o)
            o = self.norm(o)
        net_in = {"lidar": o["lidar"].float(), "target": o["target"].float(), "action_mask": obs["action_mask"].float()}
        if self.use_img:
            net_in["img"] = obs["img"].float() / 255.0
        mean = torch.clamp(self._forward(net_in).double(), -1, 1)  # ppo_agent.py:141
        std = torch.exp(self.log_std.detach().double()).expand_as(mean)
        if self.use_mask_sampling:
            action, _ = masked_discrete_actions(mean, std, obs["action_mask"], self.actions42, self.gen)
        else:
            action = torch.clamp(mean + std * torch.randn(mean.shape, dtype=mean.dtype, device=mean.device, generator=self.gen), -1, 1)
        return action.contiguous(), (mean, std)

    def _collect_overlapped(self, n_steps):
        """collect() without a store, software-pipelined by one step: the policy side of step t + 1 (state norm, actor forward,
        masked sampling: it needs lidar / target / action mask only) runs on a second stream as soon as k_observe of step t is
        done (hope_wait_observed), next to the Reeds-Shepp kernels of step t; the plan hand-off of step t + 1, which needs their
        result, waits for both.  Same kernels, same inputs, same order per stream as the plain loop."""
        env, dev = self.env, self.env.device
        main = torch.cuda.current_stream(dev)
        if self._pol_stream is None:
            self._pol_stream = torch.cuda.Stream(device=dev)
        pol = self._pol_stream
        action, dist = self.act(self.obs)
        for t in range(n_steps):
            if self.use_planner:
                action, _ = env.planner_actions(action)
            obs, reward, done, info = env.step(action)
            self.obs = obs
            if t + 1 == n_steps:
                break
            if env.wait_observed(pol):
                with torch.cuda.stream(pol):
                    action, dist = self.act(obs)
                    for x in (action,) + tuple(dist):
                        x.record_stream(main)
                main.wait_stream(pol)      # the hand-off and the next step read the sampled action
            else:
                action, dist = self.act(obs)
        return self.obs

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
        if store is None and self.fused and self.overlap and n_steps > 0:
            return self._collect_overlapped(n_steps)
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
