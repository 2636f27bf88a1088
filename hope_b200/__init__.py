"""hope_b200 — B200-native batched ParkingEnv step (the hot path of jiamiya/HOPE's src/env).

Layout
  csrc/            CUDA kernels (sm_100a) + the C ABI declared in include/hope_b200.h
  build.py         nvcc driver (in-tree libhope_b200.so)
  capi.py          ctypes binding of the C ABI
  tables.py        host-side constant tables (ray directions, own-box offsets, dist_star)
  batched_env.py   BatchedParkingEnv: N scenes on one GPU, device-tensor and host-buffer APIs
  compat/          drop-in `env` package mirroring CarParking / CarParkingWrapper for N = 1

The CUDA library is the only compute path: importing the env without a built library, or
stepping it without a CUDA device, raises.
"""
from .capi import HopeError, load_library  # noqa: F401

__all__ = ["HopeError", "load_library"]
