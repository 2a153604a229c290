"""afft_b200: B200-native (sm_100a) implementation of AFFT's fusion-and-anticipation forward path.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/afft_b200.h``), ``_capi.py`` the ctypes binding,
``models/`` the host-side mirror of the reference ``models/`` API.
"""
__version__ = "0.1.0"
