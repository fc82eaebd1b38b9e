"""dpft_b200 — the DPFT (dprt) model hot path on B200: sm_100a CUDA kernels behind a C ABI
(include/dpft_b200.h, libdpft_b200.so) and the host-side mirror of the reference's model interface
(``dpft_b200.models.build('dprt', config)`` == ``dprt.models.build``; ``dpft_b200.msda`` == the
``MultiScaleDeformableAttention`` extension module).  No CPU fallback: ops raise without the library/GPU.
"""
__version__ = "0.1.0"
