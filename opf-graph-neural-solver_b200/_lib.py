"""ctypes binding of ``libgns_b200.so`` (C ABI declared in ``include/gns_b200.h``)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SYMBOLS = [
    "gns_plan_create", "gns_plan_destroy", "gns_plan_export", "gns_dims_supported",
    "gns_param_count", "gns_workspace_bytes", "gns_forward", "gns_forward_compact", "gns_backward",
    "gns_check_topology", "gns_check_topology_async", "gns_expand_inputs", "gns_launch_info", "gns_layout_export", "gns_adam_step", "gns_measure_ffma_flops", "gns_measure_ffma2_flops",
    "gns_last_error", "gns_version",
]


def library_path() -> str:
    return os.environ.get("GNS_LIB") or os.path.join(_HERE, "libgns_b200.so")


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (``make -C csrc``)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgns_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return library_path()


def load_library():
    """Load the C-ABI library; raises if it has not been built (no fallback path)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C opf-graph-neural-solver_b200/csrc`. This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.gns_plan_create.argtypes = [i32, i32, i32, vp, vp, vp, i32, C.POINTER(vp)]
    lib.gns_plan_create.restype = i32
    lib.gns_plan_destroy.argtypes = [vp]
    lib.gns_plan_destroy.restype = None
    lib.gns_plan_export.argtypes = [vp, C.c_char_p, vp, i32]
    lib.gns_plan_export.restype = i32
    lib.gns_dims_supported.argtypes = [i32, i32]
    lib.gns_dims_supported.restype = i32
    lib.gns_param_count.argtypes = [i32, i32, i32, i32]
    lib.gns_param_count.restype = i64
    lib.gns_workspace_bytes.argtypes = [vp, i64, i32, i32, i32, i32, i32]
    lib.gns_workspace_bytes.restype = i64
    lib.gns_forward.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32,
                                vp, vp, vp, vp, vp, i64, i32, vp]
    lib.gns_forward.restype = i32
    lib.gns_forward_compact.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32,
                                        vp, vp, vp, vp, vp, i64, vp]
    lib.gns_forward_compact.restype = i32
    lib.gns_backward.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32,
                                 vp, vp, vp, vp, vp, vp, i64, vp]
    lib.gns_backward.restype = i32
    lib.gns_check_topology.argtypes = [vp, vp, vp, i64, vp]
    lib.gns_check_topology.restype = i32
    lib.gns_check_topology_async.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.gns_check_topology_async.restype = i32
    lib.gns_expand_inputs.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp]
    lib.gns_expand_inputs.restype = i32
    lib.gns_launch_info.argtypes = [vp, i64, i32, i32, i32, i32, i32, vp]
    lib.gns_launch_info.restype = i32
    lib.gns_layout_export.argtypes = [C.c_char_p, i32, i32, i32, i32, vp, i32]
    lib.gns_layout_export.restype = i32
    lib.gns_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i64, vp]
    lib.gns_adam_step.restype = i32
    lib.gns_measure_ffma_flops.argtypes = [i32, i32]
    lib.gns_measure_ffma_flops.restype = C.c_double
    lib.gns_measure_ffma2_flops.argtypes = [i32, i32]
    lib.gns_measure_ffma2_flops.restype = C.c_double
    lib.gns_last_error.argtypes = []
    lib.gns_last_error.restype = C.c_char_p
    lib.gns_version.argtypes = []
    lib.gns_version.restype = C.c_char_p
    _LIB = lib
    return lib


def last_error() -> str:
    return load_library().gns_last_error().decode()


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
