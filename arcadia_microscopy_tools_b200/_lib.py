"""ctypes binding of ``libamt_b200.so`` (C ABI declared in ``include/amt_b200.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded, every
compute entry point raises ``AmtLibraryError`` (build it with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C arcadia_microscopy_tools_b200/csrc``).
"""

from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libamt_b200.so"

AMT_OK = 0
AMT_ERR_INVALID = -1
AMT_ERR_CUDA = -2
AMT_ERR_CAPACITY = -3
AMT_ERR_UNSUPPORTED = -4

AMT_U8, AMT_U16, AMT_I32, AMT_F64, AMT_I64 = 0, 1, 2, 3, 4
AMT_MAX_RANKS = 8
AMT_EXTEND_NEAREST, AMT_EXTEND_REFLECT = 0, 1  # amt_gaussian_axis_mode
AMT_FILTER_TENSOR_CORE, AMT_FILTER_FMA = 0, 1  # amt_fov_config.plane_filter
AMT_SEG_DECISION_EXACT, AMT_SEG_FLOAT64 = 0, 1  # amt_fov_config.seg_plane_filter
STAGE_NAMES = ("dog_exact", "dog_lo", "dog_tc_axis0", "dog_tc_axis1", "select", "map", "label_thr", "regions_thr",
               "label_given", "regions_given")  # AMT_STAGE_*
# per-FOV status bits of the executor (include/amt_b200.h)
AMT_FOV_THR_CAPACITY, AMT_FOV_GIVEN_CAPACITY, AMT_FOV_GIVEN_VALUE_RANGE = 1, 2, 4
AMT_FOV_THR_EMPTY, AMT_FOV_GIVEN_EMPTY, AMT_FOV_CONSTANT_PLANE, AMT_FOV_GIVEN_NEGATIVE = 8, 16, 32, 64
AMT_MAP_SUBCLIP, AMT_MAP_RESCALE, AMT_MAP_FILL = 1, 2, 4
AMT_ACC_BASE, AMT_ACC_PER_CHANNEL = 10, 4
AMT_TABLE_BASE, AMT_TABLE_PER_CHANNEL = 16, 5


def acc_fields(n_channels: int) -> int:
    return AMT_ACC_BASE + AMT_ACC_PER_CHANNEL * n_channels


AMT_ACC3D_BASE, AMT_TABLE3D_BASE = 16, 16


def acc3d_fields(n_channels: int) -> int:
    return AMT_ACC3D_BASE + AMT_ACC_PER_CHANNEL * n_channels


def table3d_cols(n_channels: int) -> int:
    return AMT_TABLE3D_BASE + AMT_TABLE_PER_CHANNEL * n_channels


def table_cols(n_channels: int) -> int:
    return AMT_TABLE_BASE + AMT_TABLE_PER_CHANNEL * n_channels


# table column indices (see include/amt_b200.h)
COL_LABEL, COL_AREA, COL_BBOX0, COL_CENTROID0 = 0, 1, 2, 6
COL_EIG0, COL_AXIS_MAJOR, COL_AXIS_MINOR, COL_ECC, COL_ORIENT = 8, 10, 11, 12, 13
COL_PERIMETER, COL_AREA_CONVEX = 14, 15
CH_SUM, CH_MEAN, CH_MAX, CH_MIN, CH_STD = 0, 1, 2, 3, 4


class AmtLibraryError(RuntimeError):
    """The CUDA library is missing or a call into it failed."""


class MapParams(C.Structure):
    _fields_ = [
        ("lvl", C.c_double),
        ("p1", C.c_double),
        ("p2", C.c_double),
        ("o1", C.c_double),
        ("o2", C.c_double),
        ("hist_first", C.c_double),
        ("hist_last", C.c_double),
        ("flags", C.c_int32),
        ("pad", C.c_int32),
    ]


class FovConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("n_channels", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("seg_channel", C.c_int32),
        ("chunk_fovs", C.c_int32),
        ("max_labels", C.c_int32),
        ("max_label_value", C.c_int32),
        ("quantify_given_mask", C.c_int32),
        ("with_shape", C.c_int32),
        ("given_label_dtype", C.c_int32),
        ("exact_all_channels", C.c_int32),
        ("low_sigma", C.c_double),
        ("high_sigma", C.c_double),
        ("bg_percentile", C.c_double),
        ("pct_lo", C.c_double),
        ("pct_hi", C.c_double),
        ("out_lo", C.c_double),
        ("out_hi", C.c_double),
        ("plane_filter", C.c_int32),
        ("seg_plane_filter", C.c_int32),
    ]


_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/amt_b200.h one to one
SIGNATURES: dict[str, tuple] = {
    "amt_version": (_i, []),
    "amt_strerror": (C.c_char_p, [_i]),
    "amt_last_cuda_error": (C.c_char_p, []),
    "amt_launch_count": (C.c_uint64, []),
    "amt_fp64_probe": (_i, [_i, _p, C.POINTER(C.c_uint64), _p]),
    "amt_tune": (_i, [C.c_char_p, _i]),
    "amt_host_alloc": (_i, [_sz, _i, C.POINTER(_p)]),
    "amt_host_free": (_i, [_p]),
    "amt_selftest_div": (_i, [_p, _p, _i64, _p, _p]),
    "amt_gaussian_axis": (_i, [_p, _i, _d, _p, _i64, _i64, _i64, _p, _i, _p]),
    "amt_gaussian_axis_mode": (_i, [_p, _i, _d, _p, _i64, _i64, _i64, _p, _i, _i, _p]),
    "amt_dog2d": (_i, [_p, _i, _d, _p, _i64, _i64, _i64, _p, _i, _p, _i, _p, _p, _p, _p]),
    "amt_dog2d_axis0": (_i, [_p, _i, _d, _i64, _i64, _i64, _p, _i, _p, _i, _p, _p, _p]),
    "amt_dog2d_axis1": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _i, _p, _i, _p, _p]),
    "amt_tcg_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "amt_tcg_destroy": (None, [_p]),
    "amt_tcg_weights": (_i, [_p, _p, C.POINTER(_i)]),
    "amt_tcg_supported": (_i, [_i64, _i64, _i]),
    "amt_tcg_digit_bytes": (_sz, [_i64, _i64, _i64]),
    "amt_tcg_axis0": (_i, [_p, _p, _i64, _i64, _i64, _p, _i, _i, _p]),
    "amt_tcg_axis1": (_i, [_p, _p, _p, _d, _p, _i64, _i64, _i64, _p, _p, _i, _i, _p]),
    "amt_tcg_axis1_dog": (_i, [_p, _p, _p, _p, _i, _d, _p, _i64, _i64, _i64, _p, _p, _i, _i, _p]),
    "amt_gauss_lo2d": (_i, [_p, _d, _p, _i64, _i64, _i64, _p, _i, _i, _i, _p]),
    "amt_minmax_filter_axis": (_i, [_p, _i, _p, _p, _i64, _i64, _i64, _i, _i, _i, _p]),
    "amt_deinterleave_u16": (_i, [_p, _p, _i64, _i64, _i, _p]),
    "amt_sub_f64": (_i, [_p, _p, _p, _i64, _p]),
    "amt_minmax_f64": (_i, [_p, _i64, _i64, _p, _p]),
    "amt_minmax_u16": (_i, [_p, _i64, _i64, _p, _p]),
    "amt_minmax_decode": (_i, [_p, _i, _i64, _p, _p]),
    "amt_select_f64_scratch_bytes": (_sz, [_i64, _i64]),
    "amt_select_f64": (_i, [_p, _i64, _i64, C.POINTER(_i64), _i, _p, _p, _p, _sz, _p]),
    "amt_bucket12": (_i, [_p, _p, _i64, _p]),
    "amt_select_f64_bucketed": (_i, [_p, _p, _i64, _i64, C.POINTER(_i64), _i, _p, _p, _p, _sz, _p]),
    "amt_select_u16_scratch_bytes": (_sz, [_i64]),
    "amt_select_u16": (_i, [_p, _i64, _i64, C.POINTER(_i64), _i, _p, _p, _sz, _p]),
    "amt_map": (_i, [_p, _i, _p, _i64, _i64, _p, _p, _p]),
    "amt_plan_dog_rescale": (_i, [_p, _p, _i64, _d, _d, _d, _d, _d, _p, _p]),
    "amt_hist256_f64": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "amt_hist_u16": (_i, [_p, _i64, _i64, _p, _p]),
    "amt_otsu_scratch_bytes": (_sz, [_i, _i64]),
    "amt_otsu": (_i, [_p, _i, _p, _p, _i64, _p, _p, _sz, _p]),
    "amt_threshold_gt": (_i, [_p, _i, _i64, _i64, _p, _p, _p]),
    "amt_window_threshold_u16": (_i, [_p, _i64, _i64, _i64, _i, _i, _i, _d, _d, _p, _p, _p]),
    "amt_threshold_gt_image": (_i, [_p, _i, _i64, _p, _d, _p, _p]),
    "amt_label_scratch_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "amt_label": (_i, [_p, _i, _p, _i64, _i64, _i64, _i64, _i, _p, _p, _p, _sz, _p]),
    "amt_region_reduce": (_i, [_p, _p, _i, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "amt_region_finalize": (_i, [_p, _p, _i, _i64, _i64, _p, _p]),
    "amt_region_reduce3d": (_i, [_p, _p, _i, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "amt_region_finalize3d": (_i, [_p, _i64, _i, _i64, _p, _p]),
    "amt_region_shape_scratch_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "amt_region_shape": (_i, [_p, _p, _i, _p, _i64, _i64, _i64, _i64, _p, _p, _sz, _p]),
    "amt_outline_squares": (_i, [_p, _i64, _i64, _p, _i64, _p, _p]),
    "amt_outline_trace_find": (_i, [_p, _i64, _i64, _i64, _p, _p]),
    "amt_outline_trace_write": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "amt_executor_create": (_i, [C.POINTER(FovConfig), C.POINTER(_d), _i, C.POINTER(_d), _i, C.POINTER(_p)]),
    "amt_window_threshold_f64_scratch_bytes": (_sz, [_i64, _i64, _i64, _i, _i]),
    "amt_window_threshold_f64": (_i, [_p, _i64, _i64, _i64, _i, _i, _i, _d, _d, _p, _p, _p, _sz, _p]),
    "amt_li_shift_f64": (_i, [_p, _i64, _d, _p, _p]),
    "amt_li_min_gap_scratch_bytes": (_sz, [_i64]),
    "amt_li_min_gap_f64": (_i, [_p, _i64, _p, _p, _sz, _p]),
    "amt_li_split_scratch_bytes": (_sz, [_i64]),
    "amt_li_split_f64": (_i, [_p, _i64, _d, _p, _p, _p, _p, _sz, _p]),
    "amt_hist_f64": (_i, [_p, _i64, _i64, _p, _i, _p, _p]),
    "amt_pairwise_sum_scratch_bytes": (_sz, [_i64, _i64]),
    "amt_pairwise_sum_f64": (_i, [_p, _i64, _i64, _p, _p, _sz, _p]),
    "amt_executor_destroy": (None, [_p]),
    "amt_executor_uses_tensor_cores": (_i, [_p]),
    "amt_executor_decision_exact": (_i, [_p]),
    "amt_executor_retry_count": (_i64, [_p]),
    "amt_executor_last_h2d_bytes": (_i64, [_p]),
    "amt_rle_encode_host": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "amt_executor_last_plain_mask_chunks": (_i64, [_p]),
    "amt_executor_last_rle_masks": (_i64, [_p]),
    "amt_tcg_error_bound": (_d, [_p]),
    "amt_executor_set_profiling": (_i, [_p, _i]),
    "amt_executor_stage_ms": (_i, [_p, _p, C.POINTER(_i64)]),
    "amt_executor_device_bytes": (_sz, [_p]),
    "amt_executor_run_device": (_i, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "amt_executor_run_host": (_i, [_p, _p, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "amt_executor_sync": (_i, [_p]),
    "amt_executor_last_ms": (C.c_float, [_p]),
}

_lock = threading.Lock()
_lib: C.CDLL | None = None


def load() -> C.CDLL:
    """Load (once) and return the library with every prototype declared."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise AmtLibraryError(
                    f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
                    "fallback. Build it with `make -C arcadia_microscopy_tools_b200/csrc`."
                )
            try:
                lib = C.CDLL(str(LIB_PATH))
            except OSError as exc:  # pragma: no cover - depends on the host
                raise AmtLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            # AMT_TUNE="key=value,key=value": kernel tuning knobs for benches / sweeps (amt_tune)
            import os

            for item in filter(None, os.environ.get("AMT_TUNE", "").split(",")):
                key, _, value = item.partition("=")
                if lib.amt_tune(key.strip().encode(), int(value)) != AMT_OK:
                    raise ValueError(f"AMT_TUNE: unknown knob or bad value {item!r}")
            _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    """Map a negative amt_status to the Python exception the reference would raise."""
    if status == AMT_OK:
        return
    lib = load()
    msg = lib.amt_strerror(status).decode()
    where = f"{what}: " if what else ""
    if status == AMT_ERR_CUDA:
        raise AmtLibraryError(f"{where}{msg} ({lib.amt_last_cuda_error().decode()})")
    if status == AMT_ERR_INVALID:
        raise ValueError(f"{where}{msg}")
    if status == AMT_ERR_CAPACITY:
        raise MemoryError(f"{where}{msg}")
    if status == AMT_ERR_UNSUPPORTED:
        raise TypeError(f"{where}{msg}")
    raise AmtLibraryError(f"{where}status {status}")
