"""ctypes binding of libr2l_b200.so (the C ABI declared in include/r2l_b200.h).

The library is built in-tree by `make -C r2l_b200/csrc` (see __graft_entry__.build()).  There is no
fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("R2L_LIB_OVERRIDE") or os.path.join(CSRC, "libr2l_b200.so")   # (override: A/B timing of an older build, tools only)

_lib = None

# name -> (restype, argtypes); mirrors include/r2l_b200.h one to one
_PROTOTYPES = {
    "r2l_last_error": (c_char_p, []),
    "r2l_abi_version": (c_int, []),
    "r2l_packed_bytes": (c_size_t, []),
    "r2l_fwd_workspace_bytes": (c_size_t, [c_int64]),
    "r2l_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p]),
    "r2l_forward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_size_t, c_int64, c_void_p]),
    "r2l_render_poses": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "r2l_bwd_workspace_bytes": (c_size_t, [c_int64]),
    "r2l_train_fwd_saved_bytes": (c_size_t, [c_int64]),
    "r2l_train_bwd_saved_bytes": (c_size_t, [c_int64]),
    "r2l_forward_train": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_int64, c_void_p]),
    "r2l_backward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_int64, c_void_p]),
    "r2l_backward_chunked": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_int64, c_void_p, c_int, c_void_p, c_int]),
    "r2l_grad_chunk_range": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "r2l_stream_wait_grad_chunk": (c_int, [c_int, c_void_p]),
    "r2l_teacher_packed_bytes": (c_size_t, []),
    "r2l_teacher_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p]),
    "r2l_teacher_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "r2l_teacher_forward_rays": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "r2l_raw2outputs": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "r2l_sample_pdf_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "r2l_sample_pdf": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "r2l_positional_embed": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "r2l_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double, c_int64,
                              c_void_p]),
    "r2l_read_ray_shards": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int]),
    "r2l_adam_hyper": (c_int, [c_double, c_double, c_double, c_int64, c_void_p]),
    "r2l_adam_schedule_dev": (c_int, [c_double, c_double, c_double, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "r2l_dp_handle_bytes": (c_size_t, []),
    "r2l_dp_create": (c_int, [c_int, c_int, c_int64, c_void_p]),
    "r2l_dp_connect": (c_int, [c_void_p]),
    "r2l_dp_grads": (c_void_p, []),
    "r2l_dp_params": (c_void_p, []),
    "r2l_dp_slice": (c_int, [c_int64, c_int64, c_void_p, c_void_p]),
    "r2l_dp_adam_step_range": (c_int, [c_void_p, c_void_p, c_double, c_double, c_double, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "r2l_dp_adam_step": (c_int, [c_void_p, c_void_p, c_double, c_double, c_double, c_void_p, c_void_p]),
    "r2l_dp_destroy": (c_int, []),
    "r2l_debug_set_dp_grid": (c_int, [c_int, c_int]),
    "r2l_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_void_p, c_void_p]),
    "r2l_loss_scratch_bytes": (c_size_t, []),
    "r2l_mse_loss_grad": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r2l_pool_draw": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r2l_pool_update": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r2l_pool_slot_host": (c_int64, [c_int64, c_int64, ctypes.c_uint64, c_int64]),
    "r2l_debug_set_stats": (c_int, [c_void_p]),
    "r2l_set_pair_mode": (c_int, [c_int]),
    "r2l_debug_launch_count": (ctypes.c_longlong, [c_int]),
    "r2l_debug_set_accum_debias": (c_int, [c_float, c_float]),
    "r2l_set_deterministic": (c_int, [c_int]),
    "r2l_debug_set_dw_schedule": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "r2l_debug_set_trace": (c_int, [c_void_p]),
    "r2l_debug_mma_rate": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "r2l_selftest_layer": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
}


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libr2l_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def exported_symbols():
    return list(_PROTOTYPES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(r2l_b200 has no CPU or PyTorch fallback for the hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            if os.environ.get("R2L_LIB_OVERRIDE") and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().r2l_last_error()
        raise RuntimeError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
