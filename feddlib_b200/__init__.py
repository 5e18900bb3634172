"""feddlib_b200 -- B200-native finite-element assembly engine for FEDDLib's hot path.

Product code: libfeddb200.so (CUDA, sm_100a; C ABI in include/feddb200.h) plus the host-side mirror
of the reference's FE interface.  Nothing here imports oracle/ and there is no CPU fallback.
"""
from ._lib import (BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, SCATTER_ATOMIC, SCATTER_COLOURED, SCATTER_GATHER,
                   EngineRuntimeError, LogicError)
from .engine import Context, Mesh, Pattern, assemble_div_divT, assemble_div_divT_d
from .fe import FE, Domain, Map, Matrix

__all__ = ["Context", "Mesh", "Pattern", "FE", "Domain", "Map", "Matrix", "LogicError", "EngineRuntimeError",
           "assemble_div_divT", "assemble_div_divT_d", "BLOCK_SCALAR", "BLOCK_DIAG", "BLOCK_FULL",
           "SCATTER_ATOMIC", "SCATTER_COLOURED", "SCATTER_GATHER"]
