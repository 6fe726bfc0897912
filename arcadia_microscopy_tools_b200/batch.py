"""Batch path: whole fields of view through preprocess -> threshold/label -> quantify in one call.

Python face of the native ``amt_executor`` (``csrc/executor.cu``).  One executor per GPU; fields
of view are independent, so multi-GPU runs shard FOVs across processes with no collective
(SURVEY.md 8e).  The per-FOV workload is the reference call chain

    P_c  = rescale_by_percentile(subtract_background_dog(x_c, low, high, bg_pct), (p_lo, p_hi), out)
    m    = apply_threshold(P_seg, "otsu")
    SegmentationMask(m, {ch: x_c}, remove_edge_cells=True).cell_properties
    SegmentationMask(given_labels, {ch: x_c}, remove_edge_cells=True).cell_properties

(ref: ``operations.py:10-97, 135-216``; ``masks.py:38-65, 247-328``).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _gpu, _lib
from ._lib import check


@dataclass
class FovPipelineConfig:
    n_channels: int = 4
    height: int = 2048
    width: int = 2048
    seg_channel: int = 1
    chunk_fovs: int = 8
    max_labels: int = 4096
    max_label_value: int = 65535
    quantify_given_mask: bool = True
    with_shape: bool = False  # also perimeter / area_convex columns (not part of workload W)
    # host label masks of run_host: np.int64 (what the reference hands SegmentationMask, ref: model.py:215,
    # masks.py:138; copied as int64 and narrowed on the device), np.int32, or np.uint16 (Cellpose's own dtype)
    given_label_dtype: type = np.int32
    # False: only the segmentation channel (whose plane decides the labels) keeps scipy's exact operation order; the
    # other channels, which yield float planes only, take `plane_filter`.
    # True: every preprocessed plane is bit-identical to the reference's.
    exact_all_channels: bool = False
    # "tensor_core": the sigma_high Gaussian of the non-thresholded channels runs as integer Toeplitz products on
    # tcgen05 (planes equal to scipy's to ~1e-10 of the [0, 1] scale; falls back to "fma" for shapes it does not take);
    # "fma": float64 with fused multiply-adds (~1e-15).
    plane_filter: str = "tensor_core"
    # The thresholded channel when the tensor-core path is on.  "decision_exact" (default): tensor-core filter too;
    # every decision derived from its plane is taken on exactly re-evaluated samples wherever the filter's proven error
    # bound could change it, so thresholds, labels, counts and tables stay bit-identical to the reference's (its float
    # plane is within the bound).  "float64": scipy's exact operation order for that channel (plane bit-identical too).
    seg_plane_filter: str = "decision_exact"
    low_sigma: float = 0.6
    high_sigma: float = 16.0
    bg_percentile: float = 0.0
    percentile_range: tuple[float, float] = (1.0, 99.0)
    out_range: tuple[float, float] = (0.0, 1.0)


# column names of the float64 table (include/amt_b200.h)
def table_columns(channel_names: list[str]) -> list[str]:
    cols = ["label", "area", "bbox-0", "bbox-1", "bbox-2", "bbox-3", "centroid_y", "centroid_x",
            "inertia_tensor_eigvals-0", "inertia_tensor_eigvals-1", "axis_major_length", "axis_minor_length",
            "eccentricity", "orientation", "perimeter", "area_convex"]
    for name in channel_names:
        low = name.lower()
        cols += [f"intensity_sum_{low}", f"intensity_mean_{low}", f"intensity_max_{low}", f"intensity_min_{low}",
                 f"intensity_std_{low}"]
    return cols


class FovCapacityError(RuntimeError):
    """A field of view exceeded a capacity of the executor; ``.fovs`` maps FOV index -> status bits."""

    def __init__(self, message: str, fovs: dict[int, int]) -> None:
        super().__init__(message)
        self.fovs = fovs


_FATAL = (_lib.AMT_FOV_THR_CAPACITY | _lib.AMT_FOV_GIVEN_CAPACITY | _lib.AMT_FOV_GIVEN_VALUE_RANGE |
          _lib.AMT_FOV_GIVEN_NEGATIVE)


def describe_status(bits: int) -> list[str]:
    names = {_lib.AMT_FOV_THR_CAPACITY: "threshold mask has more cells than max_labels",
             _lib.AMT_FOV_GIVEN_CAPACITY: "given mask has more cells than max_labels",
             _lib.AMT_FOV_GIVEN_VALUE_RANGE: "given mask holds a value above max_label_value",
             _lib.AMT_FOV_THR_EMPTY: "no cells remain in the threshold mask after removing edge cells",
             _lib.AMT_FOV_GIVEN_EMPTY: "no cells remain in the given mask after removing edge cells",
             _lib.AMT_FOV_CONSTANT_PLANE: "the thresholded plane is constant",
             _lib.AMT_FOV_GIVEN_NEGATIVE: "given mask holds a negative value"}
    return [text for bit, text in names.items() if bits & bit]


def raise_for_status(status, config: "FovPipelineConfig") -> None:
    """Raise ``FovCapacityError`` if any FOV of a finished batch dropped data (capacity / value-range bits).
    'No cells remain' and 'constant plane' are results, not errors: the reference raises them per image
    (ref: masks.py:57-60) and a plate run must survive them (ref: model.py:276-288); they stay in ``status``."""
    status = np.asarray(status)
    bad = {int(i): int(status[i]) for i in np.flatnonzero(status & _FATAL)}
    if bad:
        first = next(iter(bad))
        raise FovCapacityError(
            f"{len(bad)} field(s) of view exceeded the executor's capacity (max_labels={config.max_labels}, "
            f"max_label_value={config.max_label_value}); first: FOV {first}: " + "; ".join(describe_status(bad[first])),
            bad)


class FovBatchExecutor:
    """Owns one native executor (streams, scratch, staging) on one GPU."""

    def __init__(self, config: FovPipelineConfig, device: int | None = None) -> None:
        dev = _gpu.require_cuda(device)
        self.config = config
        self.device = dev
        self._lib = _lib.load()
        cfg = _lib.FovConfig(
            device=dev.index, n_channels=config.n_channels, height=config.height, width=config.width,
            seg_channel=config.seg_channel, chunk_fovs=config.chunk_fovs, max_labels=config.max_labels,
            max_label_value=config.max_label_value, quantify_given_mask=1 if config.quantify_given_mask else 0,
            with_shape=1 if config.with_shape else 0,
            given_label_dtype={np.dtype(np.uint16): _lib.AMT_U16, np.dtype(np.int32): _lib.AMT_I32,
                               np.dtype(np.int64): _lib.AMT_I64}[np.dtype(config.given_label_dtype)],
            exact_all_channels=1 if config.exact_all_channels else 0,
            plane_filter={"tensor_core": _lib.AMT_FILTER_TENSOR_CORE, "fma": _lib.AMT_FILTER_FMA}[config.plane_filter],
            seg_plane_filter={"decision_exact": _lib.AMT_SEG_DECISION_EXACT, "float64": _lib.AMT_SEG_FLOAT64}[config.seg_plane_filter],
            low_sigma=config.low_sigma, high_sigma=config.high_sigma,
            bg_percentile=config.bg_percentile, pct_lo=config.percentile_range[0], pct_hi=config.percentile_range[1],
            out_lo=config.out_range[0], out_hi=config.out_range[1],
        )
        hw_lo = _gpu.gaussian_half_weights(config.low_sigma)
        hw_hi = _gpu.gaussian_half_weights(config.high_sigma)
        handle = C.c_void_p()
        check(
            self._lib.amt_executor_create(
                C.byref(cfg), hw_lo.ctypes.data_as(C.POINTER(C.c_double)), len(hw_lo) - 1,
                hw_hi.ctypes.data_as(C.POINTER(C.c_double)), len(hw_hi) - 1, C.byref(handle)),
            "amt_executor_create",
        )
        self._handle = handle
        self.n_cols = _lib.table_cols(config.n_channels)

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.amt_executor_destroy(self._handle)
            self._handle = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "FovBatchExecutor":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    @property
    def uses_tensor_cores(self) -> bool:
        """True when the non-thresholded channels' wide Gaussian runs on tcgen05 (``plane_filter`` resolved)."""
        return bool(self._lib.amt_executor_uses_tensor_cores(self._handle))

    @property
    def decision_exact(self) -> bool:
        """True when the thresholded channel runs in decision-exact mode (``seg_plane_filter`` resolved)."""
        return bool(self._lib.amt_executor_decision_exact(self._handle))

    @property
    def retry_count(self) -> int:
        """Fields of view recomputed in float64 because a candidate list of the decision-exact mode overflowed."""
        return int(self._lib.amt_executor_retry_count(self._handle))

    @property
    def last_h2d_bytes(self) -> int:
        """Host -> device bytes of the last ``run_host`` batch (label masks cross PCIe as per-row runs of equal value)."""
        return int(self._lib.amt_executor_last_h2d_bytes(self._handle))

    @property
    def last_rle_masks(self) -> int:
        """Label masks of the last ``run_host`` batch that crossed PCIe as runs (the others crossed plain: the executor
        balances host encoding time against PCIe time per chunk)."""
        return int(self._lib.amt_executor_last_rle_masks(self._handle))

    @property
    def last_plain_mask_chunks(self) -> int:
        """Chunks of the last ``run_host`` batch whose label masks were too ragged for run-length staging."""
        return int(self._lib.amt_executor_last_plain_mask_chunks(self._handle))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.amt_executor_device_bytes(self._handle))

    # ------------------------------------------------------------------ device-resident batch
    def alloc_outputs(self, n_fov: int, labels: bool = False, preprocessed: bool = False) -> dict:
        torch = _gpu.torch_mod()
        c = self.config
        dev = self.device
        out = {
            "tables_thr": torch.empty((n_fov, self.n_cols, c.max_labels), dtype=torch.float64, device=dev),
            "counts_thr": torch.empty(n_fov, dtype=torch.int32, device=dev),
            "tables_given": torch.empty((n_fov, self.n_cols, c.max_labels), dtype=torch.float64, device=dev),
            "counts_given": torch.empty(n_fov, dtype=torch.int32, device=dev),
            "thresholds": torch.empty(n_fov, dtype=torch.float64, device=dev),
            "status": torch.zeros(n_fov, dtype=torch.int32, device=dev),
            "labels_thr": None, "labels_given": None, "preprocessed": None,
        }
        if labels:
            out["labels_thr"] = torch.empty((n_fov, c.height, c.width), dtype=torch.int32, device=dev)
            out["labels_given"] = torch.empty((n_fov, c.height, c.width), dtype=torch.int32, device=dev)
        if preprocessed:
            out["preprocessed"] = torch.empty((n_fov, c.n_channels, c.height, c.width), dtype=torch.float64, device=dev)
        return out

    def run_device(self, fovs, given_labels, outputs: dict, sync: bool = True) -> float:
        """fovs: CUDA (n_fov, C, H, W) uint16 bits; given_labels: CUDA (n_fov, H, W) int32 or None.
        Work is queued on the executor's own streams (inputs must be complete: this method
        synchronises the caller's current stream first).  Returns device milliseconds when
        ``sync`` (CUDA events on the compute stream)."""
        torch = _gpu.torch_mod()
        n_fov = fovs.shape[0]
        torch.cuda.current_stream(self.device).synchronize()
        p = _gpu.ptr
        check(
            self._lib.amt_executor_run_device(
                self._handle, p(fovs), p(given_labels), n_fov, p(outputs["tables_thr"]), p(outputs["counts_thr"]),
                p(outputs["tables_given"]), p(outputs["counts_given"]), p(outputs["thresholds"]),
                p(outputs["labels_thr"]), p(outputs["labels_given"]), p(outputs["preprocessed"]),
                p(outputs.get("status"))),
            "amt_executor_run_device",
        )
        if not sync:
            return float("nan")
        check(self._lib.amt_executor_sync(self._handle), "amt_executor_sync")
        return float(self._lib.amt_executor_last_ms(self._handle))

    def set_profiling(self, enable: bool) -> None:
        """Per-stage CUDA-event timing of the runs that follow (``stage_ms``); zeroes the counters."""
        check(self._lib.amt_executor_set_profiling(self._handle, 1 if enable else 0), "amt_executor_set_profiling")

    def stage_ms(self) -> tuple[dict[str, float], int]:
        """(milliseconds per stage summed over the profiled chunks, number of chunks)."""
        ms = (C.c_double * len(_lib.STAGE_NAMES))()
        n = C.c_int64(0)
        check(self._lib.amt_executor_stage_ms(self._handle, ms, C.byref(n)), "amt_executor_stage_ms")
        return {name: float(ms[i]) for i, name in enumerate(_lib.STAGE_NAMES)}, int(n.value)

    def check_status(self, outputs: dict) -> None:
        """After a synchronised ``run_device``: raise ``FovCapacityError`` if a FOV dropped data (one small D2H)."""
        if outputs.get("status") is not None:
            raise_for_status(_gpu.to_host(outputs["status"]), self.config)

    # ------------------------------------------------------------------ host-fed batch
    def run_host(self, fovs: np.ndarray, given_labels: np.ndarray | None, out: dict | None = None,
                 on_error: str = "raise") -> dict:
        """fovs: host (n_fov, C, H, W) uint16 (pinned memory overlaps copies with compute);
        given_labels: host (n_fov, H, W) of ``config.given_label_dtype`` (int64, int32 or uint16) or None.
        Returns host arrays; ``out["status"][i]`` holds the ``AMT_FOV_*`` bits of FOV i.

        A field of view that overflows a capacity (more cells than ``max_labels``, a label value above
        ``max_label_value``, a negative label) never disturbs the others of the batch.  ``on_error="raise"``
        (default) raises ``FovCapacityError`` naming those FOVs once the whole batch is back (the reference
        never drops cells silently); ``on_error="status"`` leaves the decision to the caller."""
        c = self.config
        n_fov = fovs.shape[0]
        assert fovs.dtype == np.uint16 and fovs.flags.c_contiguous
        if given_labels is not None:
            if given_labels.dtype != np.dtype(c.given_label_dtype):
                raise TypeError(f"given_labels must be {np.dtype(c.given_label_dtype)} (config.given_label_dtype), "
                                f"got {given_labels.dtype}")
            assert given_labels.flags.c_contiguous
        if out is None:
            out = self.alloc_host_outputs(n_fov)
        vp = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        check(
            self._lib.amt_executor_run_host(
                self._handle, vp(fovs), vp(given_labels), n_fov, vp(out["tables_thr"]), vp(out["counts_thr"]),
                vp(out["tables_given"]), vp(out["counts_given"]), vp(out["thresholds"]), vp(out.get("status"))),
            "amt_executor_run_host",
        )
        if on_error == "raise" and out.get("status") is not None:
            raise_for_status(out["status"], c)
        return out

    def alloc_host_outputs(self, n_fov: int, pinned: bool = True) -> dict:
        torch = _gpu.torch_mod()
        c = self.config

        def host(shape, dtype):
            t = torch.empty(shape, dtype=dtype, pin_memory=pinned)
            return t.numpy()

        return {
            "tables_thr": host((n_fov, self.n_cols, c.max_labels), torch.float64),
            "counts_thr": host((n_fov,), torch.int32),
            "tables_given": host((n_fov, self.n_cols, c.max_labels), torch.float64),
            "counts_given": host((n_fov,), torch.int32),
            "thresholds": host((n_fov,), torch.float64),
            "status": host((n_fov,), torch.int32),
        }

    # ------------------------------------------------------------------ table -> dict
    def table_to_properties(self, table: np.ndarray, count: int, channel_names: list[str]) -> dict[str, np.ndarray]:
        """One FOV's (cols, max_labels) table -> ``cell_properties``-style dict of columns."""
        cols = table_columns(channel_names)
        out: dict[str, np.ndarray] = {}
        if count > table.shape[1]:
            raise FovCapacityError(f"{count} cells but the table holds {table.shape[1]} (max_labels)", {})
        for i, name in enumerate(cols):
            col = np.ascontiguousarray(table[i, :count])
            if name == "label" or name.startswith("bbox-"):
                col = col.astype(np.int64)
            elif name.startswith("intensity_sum_"):
                col = col.astype(np.uint64)
            out[name] = col
        return out


# ---------------------------------------------------------------------------------------------- plates
def fov_record(out: dict, j: int, n_cols_used: int | None = None) -> dict:
    """One FOV's results cut out of a batch's host outputs: counts, threshold, status and the table columns that
    are filled (``min(count, max_labels)`` cells) — what crosses ranks when a plate is gathered."""
    rec = {"threshold": float(out["thresholds"][j]), "status": int(out["status"][j]) if "status" in out else 0}
    for which in ("thr", "given"):
        cnt = int(out[f"counts_{which}"][j])
        tab = out[f"tables_{which}"][j]
        rec[f"count_{which}"] = cnt
        rec[f"table_{which}"] = np.array(tab[:, : min(cnt, tab.shape[1])], copy=True)
    return rec


def run_plate(fov_source, n_fov: int, config: FovPipelineConfig, dist=None, device: int | None = None,
              batch_fovs: int = 32, process_batch=None, dst: int = 0):
    """A whole plate (BASELINE config 5: wells x FOVs), sharded by field of view across the ranks of an initialised
    ``torch.distributed`` group: FOV i belongs to rank ``i mod world`` (the reference's analogue is
    the per-axis fan-out of ``pipeline.py:139-149``).  No data-path collective: every rank feeds its own FOVs
    through its own executor (``run_host`` in batches of ``batch_fovs``), then the per-FOV records
    (``fov_record``) are gathered on ``dst`` in FOV order.  Returns the list on ``dst``, None elsewhere.

    ``fov_source(i) -> (fov uint16 (C, H, W), given labels (H, W) of config.given_label_dtype or None)``.
    ``process_batch(fovs, givens) -> host output dict`` replaces the executor (CPU tests of the sharding logic)."""
    from .sharding import gather_fov_results, shard_indices

    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    mine = shard_indices(n_fov, rank, world)
    executor = None
    if process_batch is None:
        executor = FovBatchExecutor(config, device=device)
        process_batch = lambda f, g: executor.run_host(f, g, on_error="status")  # noqa: E731
    local: dict[int, dict] = {}
    try:
        for b0 in range(0, len(mine), batch_fovs):
            idx = mine[b0 : b0 + batch_fovs]
            items = [fov_source(int(i)) for i in idx]
            fovs = np.ascontiguousarray(np.stack([it[0] for it in items]))
            givens = None if items[0][1] is None else np.ascontiguousarray(np.stack([it[1] for it in items]))
            out = process_batch(fovs, givens)
            for j, i in enumerate(idx):
                local[int(i)] = fov_record(out, j) | {"rank": rank}
    finally:
        if executor is not None:
            executor.close()
    return gather_fov_results(local, n_fov, dist, dst=dst)
