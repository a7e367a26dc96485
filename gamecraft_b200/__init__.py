"""gamecraft_b200 — B200-native chunk mesher for Gamecraft (one hot path, drop-in behind a C ABI).

  csrc/      CUDA kernels (sm_100a) + the C ABI of libgcmesh.so (include/gcmesh.h)
  mesher     host-side mirror of the reference's BuildChunk / RebuildChunks call sites (ctypes)
  slab       z-slab decomposition across GPUs + boundary-plane buffer layout
  worldgen   deterministic synthetic-world producer (inputs for tests / bench; not on the meshing path)
"""
__version__ = "0.1.0"
