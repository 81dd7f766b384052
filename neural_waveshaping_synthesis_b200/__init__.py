"""B200-native Neural Waveshaping Synthesis forward pass (hot path only).

Host side: Python mirror of the reference's module API.  Device side: hand-written
sm_100a CUDA behind a C ABI (include/nws_b200.h, csrc/).  Importing this package
does not load the CUDA library; the first forward (or `load_library()`) does and
fails loudly when it is missing — there is no CPU fallback.
"""
__version__ = "0.1.0"
