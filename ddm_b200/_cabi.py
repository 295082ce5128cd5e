"""ctypes binding of the C ABI in ``include/dddm_b200.h`` (``ddm_b200/lib/libdddm_b200.so``).

There is no fallback: if the library has not been built (``python -m ddm_b200.build``) importing
anything that computes raises ``DDDMLibraryMissing``.  The library is opened from the package
directory (in-tree), never from site-packages or a JIT cache.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_ulonglong, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdddm_b200.so")
ABI_VERSION = 1


class DDDMLibraryMissing(ImportError):
    pass


class DDDMError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{message} (status {status})")
        self.status = status


# (restype, argtypes) for every symbol declared in include/dddm_b200.h
SIGNATURES = {
    "dddm_abi_version": (c_int, []),
    "dddm_strerror": (c_char_p, [c_int]),
    "dddm_energy_workspace_bytes": (c_size_t, [c_int, c_int]),
    "dddm_energy_workspace_reset": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dddm_energy_dist_per_row": (c_size_t, [c_int]),
    "dddm_energy_fused_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int, c_int, c_float, c_float, c_void_p]),
    "dddm_energy_fused_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_float, c_float, c_void_p]),
    "dddm_energy_fused_bf16_x0f32": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                             c_int, c_int, c_float, c_float, c_void_p]),
    "dddm_energy_fused_bf16_x0f32_supported": (c_int, [c_int, c_int]),
    "dddm_energy_terms_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_float, c_void_p]),
    "dddm_energy_terms_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_float, c_void_p]),
    "dddm_energy_terms_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_float, c_void_p]),
    "dddm_energy_terms_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_float, c_void_p]),
    "dddm_scale_inplace_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "dddm_scale_inplace_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "dddm_forward_marginal_expand_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                 c_long, c_void_p]),
    "dddm_forward_marginal_expand_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                  c_long, c_void_p]),
    "dddm_forward_marginal_concat_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dddm_forward_marginal_concat_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                                  c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dddm_backbone_scratch_bytes": (c_size_t, [c_int]),
    "dddm_layer_norm_fwd_f32": (c_int, [c_void_p] * 6 + [c_long, c_int, c_float, c_void_p]),
    "dddm_layer_norm_fwd_bf16": (c_int, [c_void_p] * 6 + [c_long, c_int, c_float, c_void_p]),
    "dddm_layer_norm_bwd_f32": (c_int, [c_void_p] * 9 + [c_size_t, c_long, c_int, c_void_p]),
    "dddm_layer_norm_bwd_bf16": (c_int, [c_void_p] * 9 + [c_size_t, c_long, c_int, c_void_p]),
    "dddm_colsum_f32": (c_int, [c_void_p] * 3 + [c_size_t, c_long, c_int, c_void_p]),
    "dddm_colsum_bf16": (c_int, [c_void_p] * 3 + [c_size_t, c_long, c_int, c_void_p]),
    "dddm_row_sqnorm_f32": (c_int, [c_void_p, c_void_p, c_long, c_long, c_void_p]),
    "dddm_rbf_tc_padded_cols": (c_long, [c_long]),
    "dddm_rbf_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_long, c_void_p]),
    "dddm_rbf_tc_scratch_bytes": (c_size_t, []),
    "dddm_rbf_kernel_sum_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_long, c_long,
                                       c_float, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "dddm_rbf_scratch_bytes": (c_size_t, [c_long, c_long]),
    "dddm_rbf_kernel_sum_f32": (c_int, [c_void_p, c_long, c_void_p, c_void_p, c_long, c_long, c_float, c_long, c_int,
                                        c_void_p, c_size_t, c_void_p, c_void_p]),
    "dddm_sigmoid_weight_sum_f32": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "dddm_bridge_step_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double,
                                     c_void_p, c_void_p, c_long, c_long, c_void_p]),
    "dddm_philox_increment": (c_ulonglong, [c_long]),
    "dddm_bridge_step_philox_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                            c_ulonglong, c_ulonglong, c_ulonglong, c_long, c_long, c_void_p]),
    "dddm_bridge_step_philox_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                             c_ulonglong, c_ulonglong, c_ulonglong, c_long, c_long, c_void_p]),
    "dddm_bridge_step_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double,
                                      c_void_p, c_void_p, c_long, c_long, c_void_p]),
    "dddm_session_create": (c_void_p, [c_int, c_int, c_int, c_int, c_int]),
    "dddm_session_destroy": (None, [c_void_p]),
    "dddm_session_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p,
                                       c_void_p]),
    "dddm_session_enqueue_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                          c_void_p, c_void_p]),
    "dddm_session_wait": (c_int, [c_void_p]),
    "dddm_session_packed_layout": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dddm_host_alloc": (c_void_p, [c_size_t]),
    "dddm_host_alloc_input": (c_void_p, [c_size_t]),
    "dddm_host_free": (None, [c_void_p]),
    "dddm_last_error": (c_int, []),
    "dddm_set_tuning": (c_int, [c_char_p, c_int]),
    "dddm_get_tuning": (c_int, [c_char_p]),
    "dddm_launch_count": (c_ulonglong, []),
    "dddm_set_trace_buffer": (c_int, [c_void_p]),
    "dddm_energy_describe": (c_int, [c_int, c_int, c_int, c_int, c_char_p, c_int]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Open the in-tree CUDA library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DDDMLibraryMissing(
            f"{LIB_PATH} not found: build the sm_100a CUDA library with `python -m ddm_b200.build` "
            "(or __graft_entry__.build()). ddm_b200 has no CPU / eager fallback.")
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if L.dddm_abi_version() != ABI_VERSION:
        raise DDDMLibraryMissing(f"ABI version mismatch: library {L.dddm_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


def strerror(status: int) -> str:
    return lib().dddm_strerror(int(status)).decode()


def check(status: int) -> None:
    """Raise like the reference would: ValueError for bad shapes (training.py:57-58), RuntimeError otherwise."""
    if status == 0:
        return
    msg = strerror(status)
    if status == -2:
        raise ValueError(msg)
    raise DDDMError(status, msg)


def set_tuning(key: str, value: int) -> None:
    check(lib().dddm_set_tuning(key.encode(), int(value)))


def get_tuning(key: str) -> int:
    return lib().dddm_get_tuning(key.encode())


def launch_count() -> int:
    return int(lib().dddm_launch_count())


def describe_energy(B: int, m: int, D: int, dtype: str = "f32") -> str:
    buf = ctypes.create_string_buffer(256)
    lib().dddm_energy_describe(B, m, D, 1 if dtype == "bf16" else 0, buf, 256)
    return buf.value.decode()
