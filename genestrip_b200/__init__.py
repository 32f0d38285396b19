"""genestrip_b200 -- B200-native (sm_100a CUDA) read-matching hot path of Genestrip behind a C ABI.

The package holds only what the path needs: csrc/ (CUDA kernels + C ABI + C++ host mirror of the
reference's goal drivers), capi.py (ctypes binding of include/genestrip_b200.h) and synth.py (synthetic
inputs).  No CPU fallback exists: without the built extension and a CUDA device every compute call raises.
"""
from .build import LIB_PATH, build_native  # noqa: F401

__version__ = "0.1.0"
