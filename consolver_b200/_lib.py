"""ctypes binding of libconsolver.so (the C ABI declared in include/consolver.h).

There is NO CPU / PyTorch fallback: if the library cannot be built or loaded, importing the kernels raises.
The .so is built in-tree by `consolver_b200.build.build_library()` (nvcc, sm_100a)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libconsolver.so")

F32, F16, BF16 = 0, 1, 2
DPM_CONVERT_NONE, DPM_CONVERT_DIV, DPM_CONVERT_LIN, DPM_CONVERT_DIV_RECIP = 0, 1, 2, 3
FLAG_VPRED, FLAG_EFF_SCALE, FLAG_X_SCALE, FLAG_PDL, FLAG_CHAIN, FLAG_LOWP_COMBINE, FLAG_X_F32, FLAG_X_WAS_LOWP = 1, 2, 4, 8, 16, 32, 64, 128
FLAG_HOST_SCALARS, FLAG_LOWP_COEF = 256, 512
POLICY_HOST_DIV, POLICY_ACT_F16, POLICY_ACT_BF16, POLICY_COEF_F16, POLICY_COEF_BF16 = 1, 2, 4, 8, 16
ABI_VERSION = 8          # must equal CONSOLVER_ABI_VERSION of include/consolver.h (tests/test_abi_cpu.py checks)
MAX_ORDER, MAX_HIDDEN, MAX_LOGITS, MAX_IN = 8, 1024, 4096, 16

_p, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes): must list every symbol declared in include/consolver.h
SIGNATURES = {
    "consolver_abi_version": (_i, []),
    "consolver_abi_hash": (C.c_uint64, []),
    "consolver_error_string": (C.c_char_p, [_i]),
    "consolver_policy_f32": (_i, [_p] * 7 + [_f] * 4 + [_p, _i] + [_p, _p] + [_i] * 8 + [_p] * 7 + [_p]),
    "consolver_step_sd": (_i, [_i, _p, _p, _f, _p, _p, _i, _p, _p, _p, _i64, _p, _i, _i, _f, _f, _f, _f, _i, _i, _i64, _p]),
    "consolver_step_fm": (_i, [_i, _i, _p, _p, _p, _i, _p, _p, _p, _i64, _p, _i, _i, _f, _i, _i, _i64, _p]),
    "consolver_step_fm_strided": (_i, [_i, _i, _p, _i64, _p, _p, _i, _p, _p, _p, _i64, _p, _i, _i, _f, _i, _i, _i64, _p]),
    "consolver_step_dpm": (_i, [_i, _i, _p, _p, _f, _p, _p, _p, _p, _p, _p, _i64, _i, _f, _f, _p, _i, _i64, _p]),
    "consolver_policy_table_f32": (_i, [_p] * 7 + [_i, _f, _f, _i, _i, _i, _i, _p, _p]),
    "consolver_policy_sample_f32": (_i, [_p] * 4 + [_p, _p] + [_i] * 7 + [_p] * 6 + [_p]),
    "consolver_policy_gauss_f32": (_i, [_p] * 7 + [_f] * 3 + [_p, _p, _p] + [_i] * 7 + [_p] * 7 + [_p]),
    "consolver_rng_state_advance": (_i, [_p, C.c_uint64, _p]),
    "consolver_torch_philox_plan": (_i, [_i64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "consolver_sd_policy_and_step": (_i, [_p] * 8 + [_f] * 4 + [_p, _p, _p] + [_i] * 5 + [_p] * 7 +
                                     [_i, _p, _p, _f, _p, _p, _i, _p, _p, _p, _i64, _i, _f, _f, _f, _f, _i, _i, _i64, _p]),
    "consolver_set_step_launch": (_i, [_i, _i]),
    "consolver_ppo_workspace": (C.c_size_t, [_i, _i, _i, _i]),
    "consolver_ppo_loss_grad_f32": (_i, [_p] * 7 + [_i, _f, _f, _i, _i, _i] + [_p] * 3 + [_i, _f, _f] + [_p] * 3 + [_p]),
    "consolver_ppo_loss_grad_allreduce_f32": (_i, [_p] * 7 + [_i, _f, _f, _i, _i, _i] + [_p] * 3 + [_i, _f, _f] + [_p] * 3 +
                                              [_p, _p]),
    "consolver_ppo_exchange_pad_words": (C.c_int64, [_i, _i, _i, _i]),
    "consolver_cosine_features_workspace": (C.c_size_t, [_i, _i]),
    "consolver_cosine_features": (_i, [_i, _p, _p, _f, _p, _i, _i, _i, _i64, _p, _p, _p]),
}

class Rng(C.Structure):
    """consolver_rng_t"""
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64), ("state", C.c_void_p), ("nthreads", C.c_uint32)]


class Peers(C.Structure):
    """consolver_peers_t"""
    _fields_ = [("buffer_ptrs_dev", C.c_void_p), ("signal_ptrs_dev", C.c_void_p), ("rank", C.c_int), ("world", C.c_int),
                ("epoch", C.c_uint32), ("stride_floats", C.c_int64), ("pad_words", C.c_int64)]


class DpmUpdate(C.Structure):
    """consolver_dpm_update_t"""
    _fields_ = [(n, C.c_float) for n in ("cx", "a0", "a1", "rinv", "a2", "rinv1", "w", "rs")]


def philox_plan(numel: int):
    """(nthreads, generator offset increment) of torch's exponential_ launch for `numel` elements"""
    nt, inc = C.c_uint32(0), C.c_uint64(0)
    check(load().consolver_torch_philox_plan(numel, C.byref(nt), C.byref(inc)), "consolver_torch_philox_plan")
    return nt.value, inc.value


_lib = None
_lock = threading.Lock()


class ConsolverError(RuntimeError):
    pass


def bind(path: str):
    """dlopen `path`, check that it was built from THIS tree's include/consolver.h (ABI version + header hash) and
    attach the ctypes signatures.  A library built from another header revision is rejected before any entry point
    with a possibly different argument list can be called."""
    lib = C.CDLL(path)
    for name in ("consolver_abi_version", "consolver_abi_hash"):
        if not hasattr(lib, name):
            raise ConsolverError(f"{path} does not export {name}: it predates this header; rebuild it "
                                 "(`python -m consolver_b200.build --force`)")
    lib.consolver_abi_version.restype, lib.consolver_abi_version.argtypes = _i, []
    lib.consolver_abi_hash.restype, lib.consolver_abi_hash.argtypes = C.c_uint64, []
    if lib.consolver_abi_version() != ABI_VERSION:
        raise ConsolverError(f"{path}: ABI version {lib.consolver_abi_version()} != {ABI_VERSION}; rebuild it")
    from .build import HEADER, header_hash

    if os.path.exists(HEADER) and lib.consolver_abi_hash() != header_hash():
        raise ConsolverError(f"{path} was built from a different include/consolver.h "
                             f"(hash {lib.consolver_abi_hash():#x} != {header_hash():#x}); rebuild it")
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ConsolverError(f"{path} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


def load(autobuild: bool = True):
    """Load libconsolver.so, rebuilding it first when its content stamp says it is not current.  A failed rebuild
    raises: a stale library is never used in its place.  CONSOLVER_NO_AUTOBUILD=1 skips the rebuild (the header-hash
    check in `bind` still applies)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if autobuild and os.environ.get("CONSOLVER_NO_AUTOBUILD") != "1":
            from .build import build_library

            try:
                build_library()
            except Exception as e:  # noqa: BLE001
                raise ConsolverError(
                    f"libconsolver.so is missing or out of date and could not be rebuilt ({e}). consolver_b200 has no "
                    "CPU/PyTorch fallback and does not fall back to a stale library: run "
                    "`python -m consolver_b200.build` with nvcc available.") from e
        if not os.path.exists(LIB_PATH):
            raise ConsolverError(f"{LIB_PATH} not found; run `python -m consolver_b200.build`")
        _lib = bind(LIB_PATH)
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().consolver_error_string(rc).decode()
        raise ConsolverError(f"{what}: {msg} (code {rc})")


def dtype_code(torch_dtype) -> int:
    import torch

    try:
        return {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}[torch_dtype]
    except KeyError:
        raise TypeError(f"consolver_b200: unsupported latent dtype {torch_dtype}") from None


def ptr_array(ptrs):
    """host array of device pointers for the `hist` arguments"""
    n = max(len(ptrs), 1)
    arr = (C.c_void_p * n)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
