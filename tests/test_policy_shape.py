"""The policy restatement used where the reference tree is not importable (hope_b200.rollout.ReferenceShapedActor) against
the reference's own `MultiObsEmbedding(ACTOR_CONFIGS)` (network.py:34-196): same parameter names and shapes, and — with the
reference's weights loaded — the same outputs.  Build container only (needs /root/reference)."""
import importlib
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref_modules():
    added = [REF_SRC, os.path.join(ROOT, "oracle", "refshim")]
    for p in added:
        sys.path.insert(0, p)
    for m in [k for k in sys.modules if k == "configs" or k == "model" or k.startswith("model.")]:
        del sys.modules[m]
    network, configs = importlib.import_module("model.network"), importlib.import_module("configs")
    assert network.__file__.startswith(REF_SRC)
    yield network, configs
    for p in added:
        sys.path.remove(p)
    for m in [k for k in sys.modules if k == "configs" or k == "model" or k.startswith("model.") or k.startswith("shapely")]:
        del sys.modules[m]
    from hope_b200 import refconfig
    refconfig._CACHE.clear()


@pytest.mark.parametrize("use_img", [False, True])
def test_restated_actor_is_the_reference_actor(ref_modules, use_img):
    from hope_b200 import rollout
    network, configs = ref_modules
    cfg = dict(configs.ACTOR_CONFIGS)
    cfg["img_shape"] = (3, 64, 64) if use_img else None
    cfg["n_modal"] = 3 + int(use_img)
    torch.manual_seed(0)
    ref = network.MultiObsEmbedding(cfg).eval()
    mine = rollout.ReferenceShapedActor(use_img=use_img).eval()
    ref_sd = ref.state_dict()
    my_sd = {k: v for k, v in mine.state_dict().items() if k != "log_std"}
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(v.shape) for k, v in my_sd.items()}
    if use_img:
        assert sum(v.numel() for v in ref_sd.values()) == 909778  # the 4-modal actor
    missing, unexpected = mine.load_state_dict(ref_sd, strict=False)
    assert missing == ["log_std"] and not unexpected
    n = 64
    g = torch.Generator().manual_seed(1)
    obs = {"lidar": torch.rand(n, 120, generator=g) * 10, "target": torch.randn(n, 5, generator=g), "action_mask": torch.rand(n, 42, generator=g)}
    if use_img:
        obs["img"] = torch.rand(n, 3, 64, 64, generator=g)
    with torch.no_grad():
        want, got = ref(obs), mine(obs)
    assert want.shape == got.shape == (n, 2)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=1e-5)
    if use_img:  # the split the conv kernel uses (FusedImgConv): conv stack first, the rest of the network from its features
        with torch.no_grad():
            conv = mine.embed_img.net[:3](obs["img"])
            assert conv.shape == (n, 2048)
            split = mine.forward_from_img_features(obs, conv)
        assert torch.equal(split, got)


def test_reference_actor_prefers_the_real_class(ref_modules):
    from hope_b200 import rollout
    net, what = rollout.reference_actor(use_img=True)
    assert type(net).__name__ == "MultiObsEmbedding" and "reference tree" in what and hasattr(net, "log_std")
