"""ctypes binding of ``libnjode_b200.so`` (C-ABI declared in ``include/njode.h``).

The library is the product path: if it is missing or does not export the expected symbols the
import of the hot path fails loudly -- there is no CPU or eager fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "NJODE_B200_LIB", os.path.normpath(os.path.join(_HERE, "..", "lib", "libnjode_b200.so")))

ABI_VERSION = 2

# enums (include/njode.h)
ACT = {"relu": 0, "tanh": 1, "sigmoid": 2, "elu": 3, "leaky_relu": 4, "selu": 5}
SCALE = {"identity": 0, "none": 0, "tanh": 1, "sigmoid": 2}
VAR = {"direct": 0, "second_moment": 1}
IMPL = {"auto": 0, "generic": 1, "tiled": 2, "rowtile": 3, "wide": 4}
IMPL_NAME = {v: k for k, v in IMPL.items()}
HDR_TOTAL_STEPS, HDR_TOTAL_SLOTS, HDR_NUM_TILES, HDR_KMAX, HDR_WORDS = 0, 1, 2, 3, 8
ARENA_KENC, ARENA_PERM, ARENA_TILE_KMAX, ARENA_TILE_SLOT_OFF, ARENA_HEADER, ARENA_KNOTS, ARENA_WORDS = 0, 1, 2, 3, 4, 5, 8
EINVAL, ECUDA, EWORKSPACE, ECAPACITY = -1, -2, -3, -4


class NjodeDesc(C.Structure):
    _fields_ = [("d_x", C.c_int32), ("d_y", C.c_int32), ("hidden", C.c_int32),
                ("n_hidden_layers", C.c_int32), ("num_moments", C.c_int32),
                ("shared_network", C.c_int32), ("activation", C.c_int32),
                ("input_scaling", C.c_int32), ("has_dt", C.c_int32), ("dt", C.c_float),
                ("impl", C.c_int32), ("reserved", C.c_int32)]


class NjodeLossDesc(C.Structure):
    _fields_ = [("ignore_first_continuity", C.c_int32), ("variance_method", C.c_int32),
                ("eps", C.c_float), ("w0", C.c_float), ("w1", C.c_float), ("reserved", C.c_int32)]


_P = C.c_void_p
_DESC = C.POINTER(NjodeDesc)
_LDESC = C.POINTER(NjodeLossDesc)
_I64, _I32, _SZ, _F = C.c_int64, C.c_int32, C.c_size_t, C.c_float

# name -> (restype, argtypes); must list every symbol include/njode.h declares
SIGNATURES = {
    "njode_abi_version": (_I32, []),
    "njode_last_error": (C.c_char_p, []),
    "njode_params_per_stack": (_I64, [_DESC]),
    "njode_param_count": (_I64, [_DESC]),
    "njode_num_stacks": (_I32, [_DESC]),
    "njode_tile_rows": (_I32, [_DESC]),
    "njode_selected_impl": (_I32, [_DESC]),
    "njode_schedule_workspace_bytes": (_SZ, [_I64, _I64, _I32]),
    "njode_schedule_build": (C.c_int, [_DESC, _P, _P, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "njode_schedule_knots": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I32, _DESC, _P, _P]),
    "njode_ckpt_row_floats": (_I64, [_DESC]),
    "njode_num_tiles": (_I64, [_DESC, _I64]),
    "njode_forward_workspace_bytes": (_SZ, [_DESC]),
    "njode_forward": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _I64, _I64, _I32,
                                _P, _P, _P, _P, _SZ, _P]),
    "njode_batch_arena_bytes": (_SZ, [_DESC, _I64, _I64, _I64, _P]),
    "njode_batch_scratch_bytes": (_SZ, [_DESC, _I64, _I64]),
    "njode_forward_batch": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I64, _P, _SZ, _I32, _P, _I64, _P, _SZ, _P, _P, _P, _P]),
    "njode_forward_batch_begin": (C.c_int, [_DESC, _P, _P, _I64, _I64, _P, _SZ, _P, _SZ, _P, _P]),
    "njode_forward_batch_finish": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I64, _P, _SZ, _I32, _P, _I64, _P, _SZ, _P, _P, _P, _P]),
    "njode_dense_workspace_bytes": (_SZ, [_DESC]),
    "njode_dense_forward": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I64, _P, _I64, _P, _P, _SZ, _P]),
    "njode_loss_workspace_bytes": (_SZ, [_I64]),
    "njode_loss": (C.c_int, [_LDESC, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _F, _P, _P, _P, _P, _SZ, _P]),
    "njode_backward_workspace_bytes": (_SZ, [_DESC, _I64]),
    "njode_backward": (C.c_int, [_DESC, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _I64, _I64, _I32,
                                 _P, _P, _P, _P, _P, _SZ, _P]),
    "njode_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _F, _P]),
    "njode_set_kernel_timing": (C.c_int, [_I32, _P, _P]),
    "njode_ffma_peak": (C.c_int, [C.POINTER(C.c_float)]),
    "njode_device_status": (C.c_int, [C.POINTER(C.c_uint32)]),
    "njode_device_status_detail": (C.c_int, [C.POINTER(C.c_uint32)]),
    "njode_debug_cta_cycles": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "njode_debug_phase": (C.c_int, [_I32, C.POINTER(C.c_uint64), _I32]),
    "njode_kernel_launches": (_I64, [_I32]),
}

_lib = None

# Writes to a parameter buffer that bypass autograd's version counters (njode_adam_step through a raw pointer) are
# counted here, keyed by the buffer's address; the reverse sweep's guard compares the count it saw at forward time.
_GENERATION = {}


def bump_generation(addr: int) -> None:
    _GENERATION[addr] = _GENERATION.get(addr, 0) + 1


def generation(addr: int) -> int:
    return _GENERATION.get(addr, 0)


def load():
    """Load the shared library once; raise RuntimeError (never fall back) if it cannot be used."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libnjode_b200.so not found at {LIB_PATH}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or `make -C neural-jump-ode_b200/csrc`. "
            f"This package has no CPU / eager fallback for the hot path.")
    import torch  # noqa: F401  (loads torch's libcudart first so both share one CUDA runtime)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    v = lib.njode_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH} has ABI version {v}, this package needs {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().njode_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


_NULLCTX = contextlib.nullcontext()


def on_device(dev):
    """``torch.cuda.device(dev)`` when dev is not current, else nothing (the context manager costs ~8 us per use)."""
    import torch
    return _NULLCTX if dev.index is None or torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


def current_stream(dev):
    """Raw handle of torch's current stream on dev."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if dev.index is None else dev.index)
    except AttributeError:                  # private API moved: the public (slower) one
        return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
