"""ctypes binding of libmrgpu.so (include/mr_b200.h) — the only way Python code in
this repo reaches the CUDA path.  There is no CPU fallback: a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libmrgpu.so")
HEADER_PATH = os.path.join(_ROOT, "include", "mr_b200.h")

MR_OK = 0
STATUS_NAMES = {
    0: "MR_OK", 1: "MR_ERR_INVALID_ARG", 2: "MR_ERR_PARSE", 3: "MR_ERR_CUDA", 4: "MR_ERR_CLOSED",
    5: "MR_ERR_UNSUPPORTED", 6: "MR_ERR_FEATURE_MISMATCH", 7: "MR_ERR_ARITHMETIC", 8: "MR_ERR_NO_DEVICE",
    9: "MR_ERR_NOT_FOUND",
}


class MrError(RuntimeError):
    """Non-zero mr_status; what the JNI shim turns into a RuntimeException (-> HTTP 500)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


def build(force: bool = False) -> str:
    """Compile libmrgpu.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)
            if f.endswith((".cu", ".cpp", ".h", ".cuh"))] + [HEADER_PATH]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        if not os.path.exists("/usr/local/cuda/bin/nvcc"):
            if os.path.exists(LIB_PATH):
                return LIB_PATH
            raise RuntimeError("nvcc not found and libmrgpu.so is not built")
        cmd = ["make", "-C", src_dir, "-j8"] + (["-B"] if force else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libmrgpu.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function include/mr_b200.h declares (MR_API ...)."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"MR_API[^;(]*?\b(mr_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.mr_last_error.restype = C.c_char_p
        _lib.mr_version.restype = C.c_char_p
        _lib.mr_kernel_launches.restype = C.c_int64
        _lib.mr_model_is_closed.restype = C.c_int32
    return _lib


def check(status: int) -> None:
    if status != MR_OK:
        raise MrError(status, lib().mr_last_error().decode("utf-8", "replace"))


class ModelInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_features", C.c_int32), ("n_trees", C.c_int32),
                ("max_leaves", C.c_int32), ("n_chunks", C.c_int32), ("has_categorical", C.c_int32),
                ("n_internal_nodes", C.c_int64), ("device_bytes", C.c_int64)]
