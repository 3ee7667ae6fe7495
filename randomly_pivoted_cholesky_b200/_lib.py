"""ctypes binding of the C ABI in include/rpc_b200.h (librpc_b200.so, hand-written sm_100a CUDA).

There is NO fallback: if the shared library cannot be built/loaded, or no sm_100 device is present,
every entry point raises.  `build()` compiles the library in-tree with nvcc (cross-compiles without
a GPU); the built .so is git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# RPC_B200_LIB: load another build of the library (A/B measurements of kernel variants; tools/build_variant.py)
LIB_PATH = os.environ.get("RPC_B200_LIB") or os.path.join(PKG_DIR, "librpc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",           # numpy/BLAS op order in the epilogues; FMAs are written explicitly
    "-shared", "-Xcompiler", "-fPIC", "--threads", "0",
]

COL_ALIGN = 128
FEAT_ALIGN = 4
KT = 16
MAX_BLOCK = 1536          # RPC_MAX_BLOCK of include/rpc_b200.h
KERN_GAUSSIAN, KERN_LAPLACE, KERN_MATERN = 0, 1, 2


MAX_PEERS = 8


class Peers(C.Structure):
    """rpc_peers of include/rpc_b200.h: per-rank base pointers of the peer-memory windows."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("shard", C.c_int64),
                ("diag_full", C.c_void_p * MAX_PEERS), ("payload", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS)]


class KernelSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nu2", C.c_int32), ("extra_stability", C.c_int32), ("reserved", C.c_int32),
                ("bandwidth", C.c_double)]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(REPO_ROOT, "include", "rpc_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile librpc_b200.so for sm_100a with nvcc.  Returns the path of the library.  `out` / `defines` build a variant
    (e.g. defines=("RPC_ROW_INTERLEAVE=0",)) to another path for A/B measurements (RPC_B200_LIB selects it at load time)."""
    if out is not None:
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + _sources()
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
        return out
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build librpc_b200.so (and there is no CPU fallback)")
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int
_DBL = C.c_double

_SIGNATURES = {
    # name: (restype, argtypes)
    "rpc_abi_version": (_INT, []),
    "rpc_abi_sizeof": (_INT, [_INT]),
    "rpc_last_error": (C.c_char_p, []),
    "rpc_ctx_create": (_INT, [_INT, C.POINTER(_P)]),
    "rpc_ctx_destroy": (_INT, [_P]),
    "rpc_ctx_launch_count": (_I64, [_P]),
    "rpc_row_norms": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P]),
    "rpc_kernel_diag": (_INT, [_P, _P, _P, _I64, _I64]),
    "rpc_kernel_rows": (_INT, [_P, _P, C.POINTER(KernelSpec), _P, _P, _I64, _INT, _I64, _P, _P, _INT, _I64, _P, _I64]),
    "rpc_kernel_block": (_INT, [_P, _P, C.POINTER(KernelSpec), _P, _P, _INT, _INT, _I64, _P, _I64]),
    "rpc_psd_diag": (_INT, [_P, _P, _P, _I64, _I64, _P, _I64]),
    "rpc_psd_rows": (_INT, [_P, _P, _P, _I64, _I64, _P, _INT, _P, _I64, _I64]),
    "rpc_psd_block": (_INT, [_P, _P, _P, _I64, _P, _INT, _P, _I64]),
    "rpc_sample_pivots": (_INT, [_P, _P, _P, _I64, _P, _INT, _P, _P]),
    "rpc_argmax": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "rpc_gather_pivots": (_INT, [_P, _P, _P, _INT, _P, _P, _INT, _I64, _P, _P, _I64, _P, _I64, _INT, _P, _I64, _I64, _I64]),
    "rpc_compact_selected": (_INT, [_P, _P, _P, _INT, _P, _P, _P, _I64, _P, _P, _P]),
    "rpc_publish_round": (_INT, [_P, _P, _P, _P, _P, _P, _INT, _P, _DBL]),
    "rpc_gram_residual": (_INT, [_P, _P, _P, _INT, _INT, _I64, _P, _I64]),
    "rpc_rejection_cholesky": (_INT, [_P, _P, _P, _INT, _I64, _P, _INT, _INT, _DBL, _P, _I64, _P, _P]),
    "rpc_pivoted_cholesky": (_INT, [_P, _P, _P, _INT, _I64, _INT, _DBL, _DBL, _P, _I64, _P, _P]),
    "rpc_select_largest": (_INT, [_P, _P, _P, _I64, _INT, _P]),
    "rpc_build_panel": (_INT, [_P, _P, _P, _I64, _INT, _P, _INT, _P, _I64, _INT, _P, _I64, _INT]),
    "rpc_build_panel_dev": (_INT, [_P, _P, _P, _I64, _INT, _P, _P, _INT, _P, _I64, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "rpc_panel_ld": (_INT, [_INT]),
    "rpc_schur_update": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _I64, _INT, _P, _I64, _P, _I64]),
    "rpc_ctx_reserve_sms": (_INT, [_P, _INT]),
    "rpc_gather_rows": (_INT, [_P, _P, _P, _I64, _P, _P, _INT, _P, _I64, _I64]),
    "rpc_schur_update_ex": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _I64, _INT, _P, _I64, _P, _P, _I64, C.POINTER(Peers),
                                   C.c_uint64, _INT]),
    "rpc_schur_update_rows": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _I64, _INT, _P, _I64, _P, _P, _I64, _P, _I64, C.POINTER(Peers),
                                     C.c_uint64, _INT]),
    "rpc_peer_alloc": (_INT, [_P, _I64, C.POINTER(_P), _P]),
    "rpc_peer_free": (_INT, [_P, _P]),
    "rpc_peer_open": (_INT, [_P, _P, C.POINTER(_P)]),
    "rpc_peer_close": (_INT, [_P, _P]),
    "rpc_peer_wait": (_INT, [_P, _P, _P, _INT, C.c_uint64, _INT, _P]),
    "rpc_schur_update_peers": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _I64, _INT, _P, _I64, _P, _I64, C.POINTER(Peers),
                                      C.c_uint64]),
    "rpc_gather_pivots_peers": (_INT, [_P, _P, _P, _INT, _P, _P, _INT, _I64, _I64, _I64, _P, _I64, _INT, _I64, _I64, _I64, _I64,
                                       C.POINTER(Peers), C.c_uint64]),
    "rpc_kernel_rows_wsum": (_INT, [_P, _P, C.POINTER(KernelSpec), _P, _P, _I64, _INT, _I64, _P, _P, _INT, _I64, _P, _P, _INT]),
    "rpc_syrk_rows": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P, _I64]),
    "rpc_rows_dot": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P, _P]),
    "rpc_rows_tdot": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P, _P]),
    "rpc_chol_solve": (_INT, [_P, _P, _P, _I64, _INT, _P, _I64, _INT]),
    "rpc_check_padding": (_INT, [_P, _P, _P, _I64, _I64, _INT, _I64, _P, C.POINTER(C.c_int32)]),
    "rpc_probe_fp64_peak": (_INT, [_P, _P, _INT, C.POINTER(_DBL)]),
    "rpc_simple_run": (_INT, [_P, _P, _INT, C.POINTER(KernelSpec), _P, _P, _INT, _I64, _P, _I64, _I64, _P, _P, _P, _INT, _INT, _P,
                              _I64, _P, _I64, _P, _I64]),
    "rpc_simple_step": (_INT, [_P, _P, C.POINTER(KernelSpec), _P, _P, _INT, _I64, _P, _I64, _I64, _P, _INT, _P, _I64, _P,
                               _I64, _P, _I64]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load (building if stale and nvcc is present) the C-ABI library.  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
            if have_nvcc or not os.path.exists(LIB_PATH):
                build()                          # raises if nvcc is missing: there is nothing else to run
            # else: a prebuilt library travelled with the tree (file times may be newer than it): use it
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if os.environ.get("RPC_B200_LIB") and not hasattr(lib, name):
                continue                     # an older variant build under A/B measurement lacks newer entry points
            fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.rpc_abi_version() != 1:
            raise RuntimeError("librpc_b200.so ABI version mismatch")
        _lib = lib
        return lib


class RpcError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().rpc_last_error().decode(errors="replace")
        raise RpcError("%s failed (rc=%d): %s" % (what or "librpc_b200 call", rc, msg))


class Context:
    """One rpc_ctx per CUDA device (scratch + launch counter)."""

    _by_device = {}
    _extra = []

    def __init__(self, device_index: int):
        self.lib = load()
        h = _P()
        check(self.lib.rpc_ctx_create(int(device_index), C.byref(h)), "rpc_ctx_create")
        self.handle = h
        self.device_index = device_index

    @classmethod
    def get(cls, device_index: int) -> "Context":
        ctx = cls._by_device.get(device_index)
        if ctx is None:
            ctx = cls(device_index)
            cls._by_device[device_index] = ctx
        return ctx

    @classmethod
    def extra(cls, device_index: int) -> "Context":
        """A further context on the device (own scratch) for work enqueued on a second stream."""
        ctx = cls(device_index)
        cls._extra.append(ctx)
        return ctx

    def launch_count(self) -> int:
        return int(self.lib.rpc_ctx_launch_count(self.handle))


def total_launches() -> int:
    return sum(c.launch_count() for c in list(Context._by_device.values()) + Context._extra)
