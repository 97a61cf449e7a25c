"""ctypes binding of libttam.so (include/ttam.h) — the only way the package reaches the GPU.

There is no CPU fallback: every op raises if the shared library is missing or the tensors are not
on a CUDA device.  `build()` compiles csrc/*.cu in-tree for sm_100a with nvcc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = Path(os.environ["TTAM_LIB"]) if os.environ.get("TTAM_LIB") else PKG_DIR / "libttam.so"   # TTAM_LIB: A/B builds
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]

ACT = {"none": 0, None: 0, "identity": 0, "relu": 1, "gelu": 2, "tanh": 3, "selu": 4}
PREC = {"fp32": 0, "tf32": 1, "bf16": 2}
PREC_X_ROUNDED, PREC_W_ROUNDED, PREC_OUT_ROUNDED, PREC_WT = 0x100, 0x200, 0x400, 0x800
OPT = {"adamw": 0, "adam": 1, "sgd": 2}
MAX_TENSORS = 48


class TensorList(C.Structure):
    _fields_ = [
        ("count", C.c_int32), ("pad_", C.c_int32),
        ("p", C.c_void_p * MAX_TENSORS), ("g", C.c_void_p * MAX_TENSORS),
        ("m", C.c_void_p * MAX_TENSORS), ("v", C.c_void_p * MAX_TENSORS),
        ("numel", C.c_int64 * MAX_TENSORS),
    ]


class TowerDesc(C.Structure):
    _fields_ = [
        ("table", C.c_void_p), ("aug", C.c_void_p), ("table_rows", C.c_int64), ("D", C.c_int64),
        ("X", C.c_void_p), ("ldx", C.c_int64), ("F", C.c_int64),
        ("W1", C.c_void_p), ("b1", C.c_void_p), ("ldw1", C.c_int64), ("H", C.c_int64),
        ("W2", C.c_void_p), ("b2", C.c_void_p), ("G1", C.c_void_p), ("c1", C.c_void_p), ("Hg", C.c_int64),
        ("G2", C.c_void_p), ("c2", C.c_void_p),
        ("dropout_p", C.c_float), ("precision", C.c_int32), ("x_rounded", C.c_int32), ("w1_rounded", C.c_int32),
        ("seed", C.c_uint64), ("rng_base", C.c_uint64),
        ("state", C.c_void_p),
        ("bag_rowptr", C.c_void_p), ("bag_entries", C.c_void_p), ("bag_tail", C.c_void_p),
        ("bag_T", C.c_int64), ("bag_tail_start", C.c_int64), ("bag_max_nnz", C.c_int64), ("bag_wgrad", C.c_int64),
        ("bag_scratch", C.c_void_p), ("bag_scratch_bytes", C.c_int64),
        ("W2r", C.c_void_p), ("W2rT", C.c_void_p), ("G1r", C.c_void_p), ("G1rT", C.c_void_p), ("G2r", C.c_void_p), ("G2rT", C.c_void_p),
    ]


class TowerBufs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("z", "hd", "a", "pre2", "g", "t", "o", "q")]


class TowerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dpre2", "dz", "dpre1", "dhd", "dW1", "db1", "dW2", "db2", "dG1", "dc1", "dG2", "dc2")] + \
               [("accumulate", C.c_int32), ("phase", C.c_int32)]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "ttam.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libttam.so (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = PKG_DIR / "build"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    global _LIB
    _LIB = None
    return LIB_PATH


_p, _i64, _i32, _f, _u64, _d = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_double

# name -> (restype, argtypes); must list every symbol include/ttam.h declares
SIGNATURES = {
    "ttam_last_error": (C.c_char_p, []),
    "ttam_version": (C.c_int, []),
    "ttam_launch_count": (C.c_int64, []),
    "ttam_device_ok": (C.c_int, []),
    "ttam_gather_rows_f32": (C.c_int, [_p, _i64, _i64, _p, _p, _i64, _i64, _i64, _p]),
    "ttam_round_tf32": (C.c_int, [_p, _i64, _i64, _i64, _p]),
    "ttam_cast_f32_to_bf16": (C.c_int, [_p, _i64, _p, _i64, _i64, _i64, _p]),
    "ttam_advance_step": (C.c_int, [_p, _u64, _p]),
    "ttam_act_fwd": (C.c_int, [_p, _p, _i64, _i64, _i32, _f, _u64, _u64, _p, _p]),
    "ttam_act_bwd": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _f, _u64, _u64, _p, _p]),
    "ttam_linear_fwd": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i32, _f, _u64, _u64, _p, _i32, _p]),
    "ttam_linear_dgrad": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _i64, _i32, _f, _i32, _i64, _i64, _i64, _i32, _p]),
    "ttam_prepare_weights": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p]),
    "ttam_linear_wgrad_workspace_bytes": (C.c_int64, [_i64, _i64, _i64]),
    "ttam_linear_wgrad": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _i32, _p, _i64, _i32, _p]),
    "ttam_gate_fwd": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "ttam_gate_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _p]),
    "ttam_augment_fwd": (C.c_int, [_p, _p, _i64, _p, _p, _p, _i64, _i64, _p]),
    "ttam_sample_negatives": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _i32, _u64, _u64, _p, _p, _p, _p]),
    "ttam_tower_fwd": (C.c_int, [C.POINTER(TowerDesc), _p, _i64, C.POINTER(TowerBufs), _p]),
    "ttam_tower_bwd_workspace_bytes": (C.c_int64, [C.POINTER(TowerDesc), _i64]),
    "ttam_tower_bwd": (C.c_int, [C.POINTER(TowerDesc), _p, _i64, C.POINTER(TowerBufs), _p, C.POINTER(TowerGrads), _p, _i64, _p]),
    "ttam_bag_supported": (C.c_int, [_i64, _i64, _i64, _i64]),
    "ttam_bag_linear_workspace_bytes": (C.c_int64, [_i64, _i64, _i64]),
    "ttam_bag_linear_fwd": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i32, _f, _u64, _u64,
                                      _p, _i32, _p, _i64, _p]),
    "ttam_bag_linear_wgrad": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _p, _i64, _p]),
    "ttam_loss_workspace_bytes": (C.c_int64, [_i64]),
    "ttam_bag_linear_wgrad_tc_workspace_bytes": (C.c_int64, [_i64, _i64, _i64]),
    "ttam_bag_linear_wgrad_tc": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _p, _i64, _p]),
    "ttam_loss_fwd_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _p, _i64, _p]),
    "ttam_loss_aug_supported": (C.c_int, [_i64, _i64]),
    "ttam_loss_aug_fwd_bwd": (C.c_int, [_p, _p, _p, _i64, _p, _i64, _p, _p, _i32, _f, _f, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _p, _i64, _p]),
    "ttam_loss_slots_fwd_bwd": (C.c_int, [_p] * 8 + [_i64, _i64, _i64, _p, _p, _f, _f, _p, _i64, _i64, _i64, _f, _p, _i64, _p]),
    "ttam_inbatch_loss_workspace_bytes": (C.c_int64, [_i64, _i64]),
    "ttam_inbatch_loss_fwd_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _i64, _p]),
    "ttam_category_alignment_workspace_bytes": (C.c_int64, [_i64, _i64, _i64]),
    "ttam_category_alignment": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _f, _p, _p, _p, _p, _i64, _p, _i64, _p]),
    "ttam_sort_workspace_bytes": (C.c_int64, [_i64]),
    "ttam_sort_rows": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i64, _p]),
    "ttam_unique_rows": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _p]),
    "ttam_long_segments_bytes": (C.c_int64, [_i64]),
    "ttam_find_long_segments": (C.c_int, [_p, _i64, _p, _p]),
    "ttam_sparse_adam_rows": (C.c_int, [_p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _i64, _p, _i64, _p, _d, _d, _d, _d, _i64, _p, _p, _p]),
    "ttam_lazy_rows": (C.c_int, [_i32, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _i64, _p, _i64, _p,
                                 _d, _d, _d, _d, _d, _d, _i64, _p, _p, _p]),
    "ttam_lazy_catchup": (C.c_int, [_i32, _p, _p, _p, _p, _i64, _p, _i64, _p, _d, _d, _d, _d, _d, _d, _i64, _p, _p]),
    "ttam_lazy_flush": (C.c_int, [_i32, _p, _p, _p, _p, _i64, _i64, _p, _d, _d, _d, _d, _d, _d, _i64, _p, _p]),
    "ttam_dense_step": (C.c_int, [_i32, C.POINTER(TensorList), _p, _d, _d, _d, _d, _d, _d, _i64, _p, _p]),
    "ttam_slot_plan_workspace_bytes": (C.c_int64, [_i64, _i64]),
    "ttam_slot_plan": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i64, _p]),
    "ttam_slot_ids": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "ttam_slot_unpack": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _p]),
    "ttam_slot_pack": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "ttam_topk_f32_workspace_bytes": (C.c_int64, [_i64, _i64, _i64, _i64]),
    "ttam_topk_f32": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "ttam_topk_f32_after": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i64, _p]),
    "ttam_score_pairs": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _p, _p]),
    "ttam_topk_bf16_workspace_bytes": (C.c_int64, [_i64, _i64, _i64, _i64]),
    "ttam_topk_bf16": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "ttam_split_bf16x3_cols": (C.c_int64, [_i64]),
    "ttam_split_bf16x3": (C.c_int, [_p, _i64, _i64, C.c_int, _p, _p]),
    "ttam_topk_f32_tc_workspace_bytes": (C.c_int64, [_i64, _i64, _i64, _i64]),
    "ttam_topk_f32_tc": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "ttam_topk_merge": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p]),
}

_LIB = None


class TtamError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libttam.so; fail loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise TtamError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError here means the header and the library diverged
        fn.restype = res
        fn.argtypes = args
    _LIB = handle
    return handle


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().ttam_last_error().decode("utf-8", "replace")
        raise TtamError(f"{what or 'ttam call'} failed with code {code}: {msg}")
