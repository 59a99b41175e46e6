"""fluxb200 — B200-native (sm_100a) FLUX denoising hot path, drop-in for diffusion-rs' Pipeline API.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI, built into libfluxb200.so),
the ctypes binding (`lib`), the operator-level mirror of the reference's backend (`ops`) and the host-side mirror
of `Pipeline::load/forward` (`pipeline`).
"""
__version__ = "0.1.0"
