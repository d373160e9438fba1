"""ctypes binding of libpf_sm100.so (C ABI: include/pf_sm100.h).

This is the only place the shared library is loaded.  It fails loudly when the library is
missing or does not export the declared symbols -- there is no Python/CPU fallback.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PF_LIB", os.path.join(HERE, "libpf_sm100.so"))  # PF_LIB: A/B-test another build

PF_ABI_VERSION = 1
PF_ERR_FASTA_RESIDUE, PF_ERR_FASTA_RAGGED, PF_ERR_FASTA_NOHEADER = -10, -11, -12
PF_PREC_FP32, PF_PREC_BF16X3, PF_PREC_BF16, PF_PREC_FP16 = 0, 1, 2, 3
PRECISIONS = {"fp32": PF_PREC_FP32, "bf16x3": PF_PREC_BF16X3, "bf16": PF_PREC_BF16, "fp16": PF_PREC_FP16}
PF_COLSUM_FLOATS = 72
KERNEL_CLASSES = ["input", "row", "colsum", "colfin", "ffn", "head"]


class PfCfg(ctypes.Structure):
    _fields_ = [("nb_blocks", c_int32), ("nb_heads", c_int32), ("embed_dim", c_int32),
                ("ffn_mult", c_int32), ("precision", c_int32)]


REDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_size_t, c_void_p)

# every symbol include/pf_sm100.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pf_abi_version": (c_int, []),
    "pf_last_error": (c_char_p, []),
    "pf_create": (c_int, [POINTER(c_void_p), POINTER(PfCfg), POINTER(c_void_p), c_int]),
    "pf_destroy": (None, [c_void_p]),
    "pf_set_precision": (c_int, [c_void_p, c_int]),
    "pf_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int64, c_int64]),
    "pf_onehot_to_idx": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pf_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64,
                           c_void_p, c_void_p, c_size_t, c_void_p, REDUCE_FN, c_void_p]),
    "pf_forward_debug": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64,
                                 c_void_p, c_void_p, c_size_t, c_void_p, REDUCE_FN, c_void_p, c_int, c_void_p]),
    "pf_dist_to_matrix": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pf_parse_fasta": (ctypes.c_longlong, [c_char_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, POINTER(c_int32),
                                           c_void_p, c_void_p, c_int32, POINTER(c_int32)]),
    "pf_format_phylip": (ctypes.c_longlong, [c_void_p, c_int, POINTER(c_char_p), c_char_p, ctypes.c_longlong]),
    "pf_neighbor_joining": (ctypes.c_longlong, [c_void_p, c_int, POINTER(c_char_p), c_char_p, ctypes.c_longlong]),
    "pf_bme_tree": (ctypes.c_longlong, [c_void_p, c_int, POINTER(c_char_p), c_int, c_char_p, ctypes.c_longlong, c_void_p]),
    "pf_last_launch_count": (c_int, [c_void_p]),
    "pf_set_peer_exchange": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p), c_size_t]),
    "pf_peer_exchange_bytes": (c_size_t, [c_size_t]),
    "pf_device_error": (c_int, [c_void_p]),
    "pf_debug_set_dump": (c_int, [c_void_p, c_void_p]),
    "pf_profile_enable": (c_int, [c_void_p, c_int]),
    "pf_profile_read": (c_int, [c_void_p, POINTER(c_float), POINTER(c_int32)]),
}

_lib = None


class PfError(RuntimeError):
    pass


def load():
    """Load libpf_sm100.so (building it first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise PfError(
                f"{LIB_PATH} is missing and could not be built ({e}). Run `python -m phyloformer_b200.build`; "
                "phyloformer_b200 has no CPU fallback.") from e
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise PfError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.pf_abi_version() != PF_ABI_VERSION:
        raise PfError(f"ABI version mismatch: library {lib.pf_abi_version()}, binding {PF_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pf_last_error()
        raise PfError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


NULL_REDUCE = ctypes.cast(None, REDUCE_FN)
