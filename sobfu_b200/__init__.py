"""sobfu_b200 -- Blackwell-native SobolevFusion solver hot path (hand-written sm_100a CUDA behind a C ABI).

The Python layer mirrors the reference's host classes (see api.py) and binds libsobfu_b200.so with ctypes; it has
no compute of its own and no fallback when the library or the GPU is missing.
"""
from ._capi import Sobfu200Error, lib  # noqa: F401
from .api import (Affine3f, DeformationField, Intr, MarchingCubes, Params, SobFusion, Solver, TsdfVolume,  # noqa: F401
                  VectorField, computeDists, depthBilateralFilter, depthTruncation)

from .parallel import SlabSolver, slab_range  # noqa: F401,E402

__all__ = ["SlabSolver", "slab_range", "Params", "Intr", "Affine3f", "TsdfVolume", "VectorField", "DeformationField", "Solver", "MarchingCubes",
           "SobFusion", "depthBilateralFilter", "depthTruncation", "computeDists", "Sobfu200Error", "lib"]
