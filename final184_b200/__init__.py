"""final184_b200 — B200-native voxel global-illumination hot path of tobyc11/Final184.

Only what the path needs lives here:
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI (libf184.so, include/f184.h)
  api.py     ctypes binding and the host-side mirror of the reference's frame section
  scene.py   scene / camera fixtures (numpy)
  fixture/   CPU input synthesiser (G-buffer, shadow map) standing in for the Vulkan passes upstream
  dist.py    one-process-per-GPU sharding over torch.distributed / NCCL
The CPU oracle lives outside the package (oracle/) and is never imported from here.
"""
from . import scene  # noqa: F401
from .api import VoxelGI, load_library, F184Error  # noqa: F401

__all__ = ["scene", "VoxelGI", "load_library", "F184Error"]
