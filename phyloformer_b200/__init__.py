"""phyloformer_b200 -- B200-native (sm_100a) implementation of Phyloformer's inference hot path.

Drop-in surface (mirrors the reference's `phyloformer` package for this path):

    from phyloformer_b200.model import Phyloformer      # == phyloformer.model.Phyloformer
    from phyloformer_b200.data import load_alignment    # == phyloformer.data.load_alignment

`forward(x)` runs hand-written CUDA kernels from `libpf_sm100.so` through the C ABI declared
in include/pf_sm100.h.  There is no CPU path: without a B200 and the built library the call
raises.
"""
__version__ = "0.1.0"
