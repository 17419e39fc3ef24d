"""parallelnbody_b200 - B200-native (sm_100a) hot path of Milias/ParallelNbody.

The product is the C-ABI shared library ``libnbody_b200.so`` (include/nbody.h). This package is the host-side
mirror of the reference's simulation-class interface (``AOctreeSearch``,
/root/reference/Source/NBody/OctreeSearch.h:111-149) over that ABI, plus synthetic initial conditions for the
harness. There is no CPU fallback: without the built CUDA library (or without a B200) every compute call raises.
"""
from .api import (NBodyError, OctreeSearch, PARTICLE_DTYPE, METHOD_DIRECT, METHOD_BARNES_HUT, lib_path, load_library,
                  measure_fp32_peak, comm_unique_id, comm_loopback_id, sort_pairs_u64)
from . import ic  # noqa: F401

__all__ = ["NBodyError", "OctreeSearch", "PARTICLE_DTYPE", "METHOD_DIRECT", "METHOD_BARNES_HUT", "lib_path",
           "load_library", "measure_fp32_peak", "comm_unique_id", "comm_loopback_id", "sort_pairs_u64", "ic"]
