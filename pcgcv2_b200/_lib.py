"""ctypes binding of libpcgc.so (the C ABI declared in include/pcgc.h).

The library is the product: nothing here falls back to a CPU implementation.  A
missing library raises ``PcgcLibraryError`` with the build command; a failed
call raises ``PcgcError`` carrying ``pcgc_last_error()``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCGC_LIB") or os.path.join(_HERE, "lib", "libpcgc.so")     # PCGC_LIB: a tuning-variant build (tools/)
CSRC = os.path.join(_HERE, "csrc")


class PcgcLibraryError(RuntimeError):
    pass


class PcgcError(RuntimeError):
    pass


c_p = ctypes.c_void_p
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pcgc.h one to one
SIGNATURES = {
    "pcgc_version": (ctypes.c_int, []),
    "pcgc_last_error": (ctypes.c_char_p, []),
    "pcgc_launch_count": (ctypes.c_uint64, []),
    "pcgc_pack_keys": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_p, c_p]),
    "pcgc_pack_keys3": (ctypes.c_int, [c_p, c_i64, c_i32, c_i32, c_i32, c_p, c_p, c_p]),
    "pcgc_child_map_k2": (ctypes.c_int, [c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "pcgc_scale_coords": (ctypes.c_int, [c_p, c_i64, ctypes.c_float, c_p, c_p]),
    "pcgc_unpack_keys": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_p]),
    "pcgc_hash_capacity": (c_i64, [c_i64]),
    "pcgc_hash_build": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_i64, c_p, c_p]),
    "pcgc_hash_keep_flags": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_i64, c_p, c_p]),
    "pcgc_hash_contains": (ctypes.c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p]),
    "pcgc_kernel_map_k3": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_i64, c_p, c_p, c_p]),
    "pcgc_stride_down_ws_bytes": (c_sz, [c_i64]),
    "pcgc_stride_down": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "pcgc_parent_info": (ctypes.c_int, [c_p, c_p, c_i64, c_p, c_p]),
    "pcgc_kernel_map_k3_from_parent": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "pcgc_upsample_keys": (ctypes.c_int, [c_p, c_i64, c_p, c_p]),
    "pcgc_argsort_ws_bytes": (c_sz, [c_i64]),
    "pcgc_argsort_u64": (ctypes.c_int, [c_p, c_i64, ctypes.c_int, c_p, c_p, c_p, c_sz, c_p]),
    "pcgc_conv_k3_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_conv_k3_packed_floats": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k3_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k3_fwd_packed": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_conv_k3_octet_packed_floats": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k3_octet_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k3_octet_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_split_h2": (ctypes.c_int, [c_p, c_i32, c_i64, c_i32, c_p, c_i32, c_p, c_p]),
    "pcgc_join_h2": (ctypes.c_int, [c_p, c_i32, c_i64, c_i32, c_p, c_i32, c_p]),
    "pcgc_conv_k3_h2_packed_words": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k3_h2_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, ctypes.c_float, c_p, c_p]),
    "pcgc_conv_k3_h2_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32,
                                            c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k3_wide_packed_bytes": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k3_wide_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, ctypes.c_float, c_p, c_p]),
    "pcgc_conv_k3_wide_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32,
                                              c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k3_octet_tc05_packed_bytes": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k3_octet_tc05_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, ctypes.c_float, c_p, c_p]),
    "pcgc_conv_k3_octet_tc05_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32,
                                                    c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k3_octet_h2_supported": (ctypes.c_int, [c_i32, c_i32]),
    "pcgc_conv_k3_octet_h2_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32,
                                                  c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k2s2_h2_packed_words": (c_sz, [c_i32, c_i32]),
    "pcgc_conv_k2s2_h2_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, ctypes.c_float, c_p, c_p]),
    "pcgc_conv_k2s2_h2_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32,
                                              c_i32, c_p, c_p]),
    "pcgc_convT_k2s2_h2_packed_words": (c_sz, [c_i32, c_i32]),
    "pcgc_convT_k2s2_h2_pack_weights": (ctypes.c_int, [c_p, c_i32, c_i32, ctypes.c_float, c_p, c_p, c_p]),
    "pcgc_convT_k2s2_h2_fwd": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32,
                                               c_p, c_p]),
    "pcgc_conv_k1_fwd": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_conv_k2s2_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_p, c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_convT_k2s2_fwd": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_i32, c_p]),
    "pcgc_conv_h2out_supported": (ctypes.c_int, [c_i32, c_i32, c_i32]),
    "pcgc_conv_k1_fwd_h2out": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_conv_k2s2_fwd_h2out": (ctypes.c_int, [c_p, c_i32, c_p, c_p, c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32,
                                                 c_p, c_p]),
    "pcgc_convT_k2s2_fwd_h2out": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_p, c_p]),
    "pcgc_irn_ws_bytes": (c_sz, [c_i64, c_i32]),
    "pcgc_irn_fwd": (ctypes.c_int, [c_p, c_p]),
    "pcgc_conv_bwd_weight_ws_bytes": (c_sz, [c_i64, c_i32, c_i32, c_i32]),
    "pcgc_colsum_ws_bytes": (c_sz, []),
    "pcgc_conv_bwd_weight": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_i32, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_sz, c_p]),
    "pcgc_conv_k2s2_bwd": (ctypes.c_int, [c_p, c_i32, c_p, c_p, c_i64, c_p, c_i32, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_p,
                                          c_sz, c_p]),
    "pcgc_convT_k2s2_bwd": (ctypes.c_int, [c_p, c_i32, c_i64, c_p, c_i32, c_p, c_i32, c_i32, c_p, c_i32, c_p, c_p, c_sz, c_p]),
    "pcgc_colsum": (ctypes.c_int, [c_p, c_i32, c_i64, c_i32, c_p, c_p, c_sz, c_p]),
    "pcgc_topk_mask_ws_bytes": (c_sz, [c_i64]),
    "pcgc_topk_mask": (ctypes.c_int, [c_p, c_i32, c_i64, c_i64, c_p, c_p, c_sz, c_p]),
    "pcgc_prune_ws_bytes": (c_sz, [c_i64]),
    "pcgc_prune": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_i32, c_i32, c_p, c_p, c_i32, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "pcgc_eb_likelihood_fwd": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_p, c_p]),
    "pcgc_eb_likelihood_bwd": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_p, c_p, c_p, c_p]),
    "pcgc_eb_cdf_table": (ctypes.c_int, [c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p]),
    "pcgc_eb_round_minmax": (ctypes.c_int, [c_p, c_i64, c_p, c_p]),
    "pcgc_eb_symbols": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_p]),
    "pcgc_rc_encode_host": (c_i64, [c_p, c_i64, c_i32, c_p, c_i64, c_p, c_i64]),
    "pcgc_rc_decode_host": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_i64, c_p, c_i64]),
    "pcgc_rc_encode_u16_host": (c_i64, [c_p, c_i64, c_i32, c_p, c_i64, c_p, c_i64]),
    "pcgc_rc_decode_u16_host": (ctypes.c_int, [c_p, c_i64, c_i32, c_p, c_i64, c_p, c_i64]),
    "pcgc_conv_k3_ones_from_parent_fwd": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p, c_i32, c_p, c_i32, c_p, c_i32, c_i32,
                                                         c_p, c_p]),
    "pcgc_irn16_second_stage_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_p, ctypes.c_float, c_p, c_p, c_p,
                                                   c_p, c_i32, c_p, c_i32, c_p, c_i32, c_p, c_p]),
    "pcgc_conv_k3_octet_h2_k1_supported": (ctypes.c_int, [c_i32, c_i32, c_i32]),
    "pcgc_conv_k3_octet_h2_k1_fwd": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, ctypes.c_float, c_p, c_i32, c_i32, c_p, c_p, c_i32,
                                                    c_p, c_i32, c_p, c_i32, c_p, c_i32, c_p, c_p]),
    "pcgc_symbol_ranges": (ctypes.c_int, [c_p, c_i64, c_p, c_i32, c_i32, c_p, c_p, c_p]),
    "pcgc_rc_encode_ranges_host": (c_i64, [c_p, c_i64, c_p, c_i64]),
    "pcgc_bce_isin_ws_bytes": (c_sz, []),
    "pcgc_bce_isin": (ctypes.c_int, [c_p, c_i32, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "pcgc_d1_sqdist": (ctypes.c_int, [c_p, c_i64, c_p, c_i64, c_p, c_i64, c_i32, c_p, c_p, c_p]),
    "pcgc_ply_count_lines_host": (c_i64, [c_p, c_i64]),
    "pcgc_ply_parse_ascii_host": (c_i64, [c_p, c_i64, c_p, c_i64]),
    "pcgc_ply_format_ascii_host": (c_i64, [c_p, c_i64, c_p, c_i64]),
    "pcgc_octree_encode_host": (c_i64, [c_p, c_i64, c_p, c_i64]),
    "pcgc_octree_decode_host": (c_i64, [c_p, c_i64, c_p, c_i64]),
}

class IrnArgs(ctypes.Structure):
    """pcgc_irn_args (include/pcgc.h)."""
    _fields_ = [("n", c_i64), ("c", c_i32), ("reserved", c_i32), ("nbr", c_p), ("parent_nbr", c_p), ("x", c_p), ("x_h2", c_p),
                ("out", c_p), ("out_h2", c_p), ("x_ld", c_i32), ("x_h2_ld", c_i32), ("out_ld", c_i32), ("out_h2_ld", c_i32),
                ("route", c_i32 * 3), ("inv_scale", ctypes.c_float * 3), ("w3", c_p * 3), ("b3", c_p * 3), ("w1", c_p * 2),
                ("b1", c_p * 2), ("ws", c_p), ("ws_bytes", c_sz), ("overflow", c_p)]


ROUTE_H2_GATHER, ROUTE_H2_OCTET, ROUTE_TF32_GATHER, ROUTE_TF32_OCTET, ROUTE_FP32, ROUTE_WIDE = range(6)

_lib = None


def build(jobs: int = 8, verbose: bool = False) -> str:
    """Compile libpcgc.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise PcgcLibraryError("building libpcgc.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


def lib():
    """The loaded library; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcgcLibraryError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()); "
                "there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI and the binding diverge
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "pcgc"):
    if rc < 0:
        raise PcgcError(f"{what} failed ({rc}): {lib().pcgc_last_error().decode(errors='replace')}")
    return rc


def launch_count() -> int:
    return int(lib().pcgc_launch_count())
