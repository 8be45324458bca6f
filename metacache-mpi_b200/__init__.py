"""metacache-mpi_b200: B200-native build/query hot path of MetaCache-MPI.

Layout
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI (libmcgpu.so, include/mcgpu.h)
  api.py     ctypes binding of the C-ABI + host-side mirror of the reference seams
  taxonomy.py / synth.py   host logic (lineage tables, seeded synthetic data)
The directory name contains a '-', so import it through `__graft_entry__.load_package()`
(registered as module `metacache_mpi_b200`).
"""
from . import synth, taxonomy  # noqa: F401
