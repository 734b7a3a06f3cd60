"""flvis_b200 -- Blackwell-native FLVIS hot path (pyramidal LK, Shi-Tomasi + FeatureDEM selection, local BA).

The product is `libflvis_b200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/flvis_b200.h`) plus the C++ host classes in `flvis_b200/host/`.  This Python package is only
a ctypes binding used by the tests, `bench.py` and the multi-GPU driver; there is no CPU fallback:
importing `flvis_b200.capi` and calling `load_library()` raises if the shared library has not been built.
"""
__version__ = "0.1"
