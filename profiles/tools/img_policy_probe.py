import torch, time, sys
sys.path.insert(0, ".")
from hope_b200 import rollout
dev = torch.device("cuda", 0)
n = 65536
net = rollout.ReferenceShapedActor(use_img=True).to(dev).eval()
img8 = torch.randint(0, 255, (n, 3, 64, 64), dtype=torch.uint8, device=dev)
obs = {"lidar": torch.randn(n, 120, device=dev), "target": torch.randn(n, 5, device=dev), "action_mask": torch.rand(n, 42, device=dev)}
def timed(f, reps=3):
    for _ in range(2): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
f32 = torch.zeros((n, 3, 64, 64), dtype=torch.float32, device=dev)
def base():
    torch.mul(img8, 1.0 / 255.0, out=f32)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return net({**obs, "img": f32})
print("baseline (f32 image, NCHW, autocast):", timed(base), "ms")
torch.backends.cudnn.benchmark = True
print("  + cudnn.benchmark:", timed(base), "ms")
def bf16_in():
    x = img8.to(torch.bfloat16).mul_(1.0 / 255.0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return net({**obs, "img": x})
print("bf16 image:", timed(bf16_in), "ms")
def cl():
    x = img8.to(torch.bfloat16).mul_(1.0 / 255.0).contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return net({**obs, "img": x})
net_cl = net.to(memory_format=torch.channels_last)
print("bf16 + channels_last:", timed(cl), "ms")
enc = net.embed_img
def enc_only():
    x = img8.to(torch.bfloat16).mul_(1.0 / 255.0).contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return enc(x)[0]
print("image encoder alone:", timed(enc_only), "ms")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    x = img8.to(torch.bfloat16).mul_(1.0 / 255.0).contiguous(memory_format=torch.channels_last)
    for name, layer in enc.net.named_children():
        t = timed(lambda: layer(x)); x = layer(x); print("   layer", name, type(layer).__name__, tuple(x.shape), round(t, 2), "ms")
