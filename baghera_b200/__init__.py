"""baghera_b200 — B200-native hot path of BAGHERA's gene-heritability sampler.

Only what the path needs lives here: ``csrc/`` (hand-written sm_100a CUDA + the C ABI declared in
``include/baghera_b200.h``), the ctypes binding (``_lib``), the device engine (``engine``), the
batched-chain NUTS driver (``sampler``) and the host-side mirror of the reference's operator
interface (``gene_regression.analyse_normal``, ``heritability.gw_normal``, ``command``, ``cli``).
There is no CPU fallback: every compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
