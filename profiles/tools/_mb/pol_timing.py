import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))
from hope_b200 import rollout, capi
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "libpol_timing.so"))
dev = torch.device("cuda", 0)
net = rollout.ReferenceShapedActor().to(dev).eval()
n = 65536
fp = rollout.FusedPolicy(net, n, dev)
obs = {"lidar": torch.randn(n, 120, device=dev), "target": torch.randn(n, 5, device=dev), "action_mask": torch.rand(n, 42, device=dev)}
lib.hope_policy_forward.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(capi.PolicyWeights), C.c_void_p, C.c_void_p]
for _ in range(3):
    lib.hope_policy_forward(n, obs["lidar"].data_ptr(), obs["target"].data_ptr(), obs["action_mask"].data_ptr(), C.byref(fp.weights), fp.out.data_ptr(), None)
torch.cuda.synchronize()
t = (C.c_longlong * 32)()
lib.hope_policy_debug_times(t)
names = {0: "start", 1: "inputs staged", 2: "embed L1", 3: "embed L2", 4: "LN1", 5: "head loop + residual", 6: "LN2 + FF", 7: "output head GEMM", 8: "final linear"}
for k in range(1, 9):
    print(f"{names[k]:24s} {t[k] - t[k-1]:8d} cycles")
print("one head (hd=3): qkv gemm", t[11] - t[10], "store+sync", t[12] - t[11], "attention+sync", t[13] - t[12], "out gemm", t[14] - t[13])
print("total", t[8] - t[0])
