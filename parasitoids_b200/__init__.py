"""parasitoids_b200 -- B200 (sm_100a) implementation of the Parasitoids
drift-diffusion forward solve behind the reference's own Python entry points.

Modules mirror the reference files they replace:

    ParasitoidModel   flight probability, BVN cell masses, per-day kernel
    CalcSol           convolution chain (get_solutions / get_populations)
    cuda_lib          the CudaSolve backend seam
    Run               Params, main, and the fused ``solve``
    globalvars        the ``cuda`` switch (always True here)

All arithmetic runs in libpkb200.so (parasitoids_b200/csrc, C ABI in
include/pkb200.h).  There is no CPU fallback.
"""
__all__ = ['ParasitoidModel', 'CalcSol', 'cuda_lib', 'Run', 'globalvars']
