"""pttools_b200: B200-native (sm_100a CUDA) implementation of the PTtools hot path.

The batched self-similar fluid-shell solve (pttools.bubble) and the Sound Shell Model gravitational-wave spectrum
(pttools.ssmtools / pttools.omgw0) over (v_wall, alpha_n, model) sweeps, behind the reference's call surface.
All arithmetic on the path runs in hand-written CUDA kernels (libpttools_b200.so, C ABI in include/pttools_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"
