"""GPU parity of the fused batch path (the native executor, through the C ABI) against the
oracle on small seeded FOVs, against the golden config-1 fixture, and through size-independent
properties at the full 2048x2048 size."""

from __future__ import annotations

import dataclasses

import hashlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

from arcadia_microscopy_tools_b200 import _gpu  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig, table_columns  # noqa: E402
from arcadia_microscopy_tools_b200.synthetic import make_fov  # noqa: E402

NAMES = ["BRIGHTFIELD", "DAPI", "FITC", "TRITC"]
INT_PROPS = ["intensity_sum", "intensity_mean", "intensity_max", "intensity_min", "intensity_std"]
MORPH = ["label", "area", "bbox", "centroid", "axis_major_length", "axis_minor_length", "eccentricity", "orientation"]


def oracle_fov(fov, given, seg, bg, pr):
    pre = [oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[c], 0.6, 16.0, bg), pr, (0, 1)) for c in range(fov.shape[0])]
    mask = oracle.apply_threshold(pre[seg])
    res = {"pre": np.stack(pre)}
    try:
        res["labels_thr"] = oracle.process_mask(mask, True)
    except ValueError:
        res["labels_thr"] = np.zeros(mask.shape, np.int64)
    res["labels_given"] = oracle.process_mask(given.astype(np.int64), True)
    chans = {n.lower(): fov[i] for i, n in enumerate(NAMES[: fov.shape[0]])}
    for key in ("labels_thr", "labels_given"):
        if res[key].max() > 0:
            res["props_" + key[7:]] = oracle.cell_properties(res[key], chans, MORPH, INT_PROPS)
    return res


def check_table(ex, table, count, want, n_ch, ctx):
    props = ex.table_to_properties(table, count, NAMES[:n_ch])
    for key, w in want.items():
        k2 = {"centroid_y": "centroid_y", "centroid_x": "centroid_x"}.get(key, key)
        g = props[k2]
        if key in ("label", "area") or key.startswith(("bbox", "intensity_sum", "intensity_max", "intensity_min")):
            assert np.array_equal(g.astype(np.float64), w.astype(np.float64)), (ctx, key)
        else:
            atol = 1e-9 * max(1.0, float(np.abs(w).max()))
            bad = ~np.isclose(g, w, rtol=1e-5, atol=atol)
            if key == "orientation":
                # an axis direction: -pi/2 and +pi/2 are the same line (sign of a zero mu11)
                d = np.abs(g - w)
                bad = np.minimum(d, np.pi - d) > 1e-5
            if key in ("orientation", "eccentricity"):
                # round-off determined for (near-)symmetric specks in the reference itself: skip those cells
                bad &= want["area"] > 12
                bad &= np.abs(want["axis_major_length"] - want["axis_minor_length"]) > 1e-6 * want["axis_major_length"]
            assert not bad.any(), (ctx, key, np.flatnonzero(bad)[:5], g[bad][:3], w[bad][:3])


# Tolerance of the preprocessed float planes of the channels that are NOT thresholded, in the executor's default
# mode (fused multiply-adds there): 1e-12 of the plane's [0, 1] scale; the north star allows 1e-5 relative.
CONTRACTED_PLANE_ATOL = 1e-12
# ... and when their sigma = 16 Gaussian runs on the tensor cores (the default where the shape allows it: integer
# Toeplitz products with 32-bit weights, csrc/tcgauss.cu): the Gaussian itself is within 2e-10 of scipy's on the
# [0, 1] input scale (tests/test_gpu_tcgauss.py); the percentile rescale divides by p99 - p1 of the DoG plane
# (~0.03 for these images), hence 1e-8 on the rescaled plane.  Measured: ~1e-10.
TENSOR_CORE_PLANE_ATOL = 1e-8


@pytest.mark.parametrize("mode", ["tensor_core", "tensor_core_float64_seg", "fma", "exact"])
@pytest.mark.parametrize("shape,C,seg,bg", [((256, 256), 4, 1, 0.0), ((192, 320), 2, 0, 25.0), ((130, 100), 3, 2, 90.0)])
def test_executor_matches_oracle(shape, C, seg, bg, mode):
    exact_all = mode == "exact"
    n_fov = 5
    fovs, givens = [], []
    for i in range(n_fov):
        f, g, _ = make_fov(1000 + i, C, shape[0], shape[1], 40)
        fovs.append(f), givens.append(g)
    fovs, givens = np.stack(fovs), np.stack(givens)
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=seg, chunk_fovs=2, max_labels=512,
                            max_label_value=int(givens.max()), bg_percentile=bg, exact_all_channels=exact_all,
                            plane_filter="fma" if mode == "fma" else "tensor_core",
                            seg_plane_filter="float64" if mode == "tensor_core_float64_seg" else "decision_exact")
    with FovBatchExecutor(cfg) as ex:
        # the tensor-core path takes planes of at least 128 x 128 with a width that is a multiple of 16
        tc = mode.startswith("tensor_core") and shape[0] >= 128 and shape[1] >= 128 and shape[1] % 16 == 0
        dx = tc and mode == "tensor_core"  # decision-exact thresholded channel: its float plane is within the bound too
        assert ex.uses_tensor_cores == tc and ex.decision_exact == dx
        out = ex.alloc_outputs(n_fov, labels=True, preprocessed=True)
        ms = ex.run_device(_gpu.to_device(fovs), _gpu.to_device(givens), out)
        assert ms > 0
        host = {k: _gpu.to_host(v) for k, v in out.items() if v is not None}
        # the same batch through the host-fed entry point must give identical bytes
        hout = ex.run_host(fovs, givens)
    for k in ("tables_thr", "counts_thr", "tables_given", "counts_given", "thresholds"):
        a, b = host[k], hout[k]
        if k.startswith("tables"):
            for i in range(n_fov):
                cnt = int(host["counts_" + k[7:]][i])
                assert np.array_equal(a[i][:, :cnt], b[i][:, :cnt], equal_nan=True), k
        else:
            assert np.array_equal(a, b), k
    for i in range(n_fov):
        want = oracle_fov(fovs[i], givens[i], seg, bg, (1, 99))
        # the segmentation channel's plane decides the labels: bit-identical in both modes; the other planes are
        # bit-identical with exact_all_channels, within CONTRACTED_PLANE_ATOL otherwise
        if not dx:
            assert np.array_equal(host["preprocessed"][i, seg], want["pre"][seg]), f"segmentation plane differs (fov {i})"
        if exact_all:
            assert np.array_equal(host["preprocessed"][i], want["pre"]), f"preprocessed planes differ (fov {i})"
        else:
            err = np.max(np.abs(host["preprocessed"][i] - want["pre"]))
            assert err <= (TENSOR_CORE_PLANE_ATOL if tc else CONTRACTED_PLANE_ATOL), f"fov {i}: {err:.3e}"
        assert np.array_equal(host["labels_thr"][i], want["labels_thr"]), f"threshold labels differ (fov {i})"
        assert np.array_equal(host["labels_given"][i], want["labels_given"]), f"given labels differ (fov {i})"
        assert host["counts_thr"][i] == want["labels_thr"].max() and host["counts_given"][i] == want["labels_given"].max()
        if "props_thr" in want:
            check_table(ex, host["tables_thr"][i], int(host["counts_thr"][i]), want["props_thr"], C, (i, "thr"))
        check_table(ex, host["tables_given"][i], int(host["counts_given"][i]), want["props_given"], C, (i, "given"))


def test_executor_golden_config1(golden):
    fov, given = golden["fov"][None], golden["given"][None]
    for bg, exact_all, seg_filter in ((0.0, True, "float64"), (90.0, True, "float64"), (0.0, False, "float64"),
                                      (90.0, False, "float64"), (0.0, False, "decision_exact"), (90.0, False, "decision_exact")):
        cfg = FovPipelineConfig(n_channels=4, height=256, width=256, seg_channel=1, chunk_fovs=1, max_labels=256,
                                max_label_value=int(given.max()), bg_percentile=bg, exact_all_channels=exact_all,
                                seg_plane_filter=seg_filter)
        with FovBatchExecutor(cfg) as ex:
            assert ex.decision_exact == (seg_filter == "decision_exact")
            out = ex.alloc_outputs(1, labels=True, preprocessed=True)
            ex.run_device(_gpu.to_device(fov), _gpu.to_device(given), out)
            host = {k: _gpu.to_host(v) for k, v in out.items() if v is not None}
        tag = f"bg{int(bg)}"
        for c in range(4):  # default mode: only the thresholded channel's plane is bit-identical by construction
            sha = hashlib.sha256(np.ascontiguousarray(host["preprocessed"][0, c]).tobytes()).hexdigest()
            if exact_all or (c == 1 and seg_filter == "float64"):
                assert sha == str(golden[f"{tag}/pre_sha256"][c]), (tag, c)
        assert host["thresholds"][0] == float(golden[f"{tag}/threshold"])
        assert np.array_equal(host["labels_thr"][0], golden[f"{tag}/labels_thr"])
        assert np.array_equal(host["labels_given"][0], golden[f"{tag}/labels_given"])
        cols = table_columns(NAMES)
        k = int(host["counts_thr"][0])
        tab = host["tables_thr"][0][:, :k]
        assert np.array_equal(tab[cols.index("area")], golden[f"{tag}/thr/area"])
        for name in ("brightfield", "dapi", "fitc", "tritc"):
            assert np.array_equal(tab[cols.index(f"intensity_sum_{name}")], golden[f"{tag}/thr/intensity_sum_{name}"].astype(np.float64))
            assert np.allclose(tab[cols.index(f"intensity_mean_{name}")], golden[f"{tag}/thr/intensity_mean_{name}"], rtol=1e-12)


def test_executor_full_size_properties():
    """Config-2 shape (4 x 2048 x 2048): properties that do not need the (slow) oracle —
    replicated FOVs give identical tables, sum of areas == labelled pixels, per-cell intensity
    sums add up to the masked image sum, bbox contains the centroid, labels are 1..K."""
    fov, given, k = make_fov(20260000, 4, 2048, 2048, 2000)
    n_fov = 3
    fovs = np.stack([fov, fov[:, ::-1].copy(), fov])
    givens = np.stack([given, given[::-1].copy(), given])
    cfg = FovPipelineConfig(chunk_fovs=2, max_labels=4096, max_label_value=int(given.max()))
    with FovBatchExecutor(cfg) as ex:
        out = ex.alloc_outputs(n_fov, labels=True)
        ex.run_device(_gpu.to_device(fovs), _gpu.to_device(givens), out)
        host = {kk: _gpu.to_host(v) for kk, v in out.items() if v is not None}
    cols = table_columns(NAMES)
    for which in ("thr", "given"):
        counts = host[f"counts_{which}"]
        assert counts[0] == counts[2] and counts[0] > 100
        assert np.array_equal(host[f"tables_{which}"][0][:, : counts[0]], host[f"tables_{which}"][2][:, : counts[2]], equal_nan=True)
        assert counts[1] == counts[0]  # a vertical flip keeps the component count
        for i in range(n_fov):
            lab = host[f"labels_{which}"][i]
            kk = int(counts[i])
            tab = host[f"tables_{which}"][i][:, :kk]
            assert lab.max() == kk and np.array_equal(np.unique(lab), np.arange(kk + 1))
            area = tab[cols.index("area")]
            assert np.array_equal(area, np.bincount(lab.ravel(), minlength=kk + 1)[1:])
            for c, name in enumerate(NAMES):
                s = tab[cols.index(f"intensity_sum_{name.lower()}")]
                assert s.sum() == float(fovs[i, c][lab > 0].sum(dtype=np.uint64))
                assert np.array_equal(s, np.bincount(lab.ravel(), weights=fovs[i, c].ravel().astype(np.float64), minlength=kk + 1)[1:])
            cy, cx = tab[cols.index("centroid_y")], tab[cols.index("centroid_x")]
            assert np.all(cy >= tab[cols.index("bbox-0")]) and np.all(cy <= tab[cols.index("bbox-2")] - 1)
            assert np.all(cx >= tab[cols.index("bbox-1")]) and np.all(cx <= tab[cols.index("bbox-3")] - 1)
            border = np.concatenate([lab[0], lab[-1], lab[:, 0], lab[:, -1]])
            assert not border.any()  # remove_edge_cells
    # the given mask: every surviving cell keeps its generated area
    lab0 = host["labels_given"][0]
    assert np.array_equal(lab0 > 0, oracle.labeling.clear_border(given) > 0)
    # exact_all_channels only changes the float planes of the channels that are not thresholded: thresholds, labels
    # and tables are the same bytes as in the default mode
    cfg_exact = FovPipelineConfig(chunk_fovs=2, max_labels=4096, max_label_value=int(given.max()), exact_all_channels=True)
    with FovBatchExecutor(cfg_exact) as ex:
        out = ex.alloc_outputs(n_fov, labels=True)
        ex.run_device(_gpu.to_device(fovs), _gpu.to_device(givens), out)
        strict = {kk: _gpu.to_host(v) for kk, v in out.items() if v is not None}
    for key in ("thresholds", "counts_thr", "counts_given", "labels_thr", "labels_given"):
        assert np.array_equal(strict[key], host[key]), key
    for which in ("thr", "given"):
        for i in range(n_fov):
            kk = int(host[f"counts_{which}"][i])
            assert np.array_equal(strict[f"tables_{which}"][i][:, :kk], host[f"tables_{which}"][i][:, :kk], equal_nan=True)


def test_run_host_label_mask_dtypes_match():
    """Host label masks may travel as int64 (the reference's dtype at this boundary, ref: masks.py:138), int32 or
    uint16 (Cellpose's mask dtype): same tables; a dtype other than the configured one is a TypeError."""
    n_fov, C, shape = 3, 2, (128, 160)
    fovs, givens = [], []
    for i in range(n_fov):
        f, g, _ = make_fov(2000 + i, C, shape[0], shape[1], 30)
        fovs.append(f), givens.append(g)
    fovs, givens = np.stack(fovs), np.stack(givens).astype(np.int32)
    outs = []
    for dt in (np.int32, np.uint16, np.int64):
        cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=1, chunk_fovs=2, max_labels=256,
                                max_label_value=int(givens.max()), given_label_dtype=dt)
        with FovBatchExecutor(cfg) as ex:
            outs.append(ex.run_host(fovs, givens.astype(dt)))
            with pytest.raises(TypeError):
                ex.run_host(fovs, givens.astype(np.int16))
    assert outs[0]["counts_given"].min() > 0
    for other in outs[1:]:
        assert np.array_equal(outs[0]["counts_given"], other["counts_given"])
        assert not other["status"].any()
        for i in range(n_fov):
            cnt = int(outs[0]["counts_given"][i])
            assert np.array_equal(outs[0]["tables_given"][i][:, :cnt], other["tables_given"][i][:, :cnt], equal_nan=True)


def test_status_isolates_a_field_of_view_that_overflows():
    """SURVEY 5 / ref: model.py:276-288: a bad field of view must not poison its batch.  FOV 1 is speckle whose
    threshold mask has more components than max_labels; FOV 2's given mask holds a value above max_label_value and
    a negative one; FOV 3 is constant.  FOVs 0 and 4 come back exactly as when processed alone."""
    from arcadia_microscopy_tools_b200 import _lib
    from arcadia_microscopy_tools_b200.batch import FovCapacityError

    C, shape, cells = 2, (128, 160), 20
    rng = np.random.default_rng(5)
    fovs, givens = [], []
    for i in range(5):
        f, g, _ = make_fov(4000 + i, C, shape[0], shape[1], cells)
        fovs.append(f), givens.append(g.astype(np.int64))
    fovs, givens = np.stack(fovs), np.stack(givens)
    fovs[1, 0] = np.where(rng.random(shape) < 0.08, 40000, 300).astype(np.uint16)  # isolated bright pixels
    givens[2, 40, 40] = 70000
    givens[2, 41, 41] = -3
    fovs[3] = 500
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=0, chunk_fovs=2, max_labels=64,
                            max_label_value=cells + 5, given_label_dtype=np.int64)
    with FovBatchExecutor(cfg) as ex:
        with pytest.raises(FovCapacityError) as err:
            ex.run_host(fovs, givens)
        assert set(err.value.fovs) == {1, 2}
        out = ex.run_host(fovs, givens, on_error="status")
        st = out["status"]
        assert st[1] & _lib.AMT_FOV_THR_CAPACITY and out["counts_thr"][1] > cfg.max_labels
        assert st[2] & _lib.AMT_FOV_GIVEN_VALUE_RANGE and st[2] & _lib.AMT_FOV_GIVEN_NEGATIVE
        assert st[3] & _lib.AMT_FOV_CONSTANT_PLANE and st[3] & _lib.AMT_FOV_THR_EMPTY and out["counts_thr"][3] == 0
        assert st[0] == 0 and st[4] == 0
        with pytest.raises(FovCapacityError):
            ex.table_to_properties(out["tables_thr"][1], int(out["counts_thr"][1]), NAMES[:C])
        for i in (0, 4):
            alone = ex.run_host(fovs[i : i + 1], givens[i : i + 1])
            for which in ("thr", "given"):
                k = int(alone[f"counts_{which}"][0])
                assert k == out[f"counts_{which}"][i] and k > 0
                assert np.array_equal(alone[f"tables_{which}"][0][:, :k], out[f"tables_{which}"][i][:, :k], equal_nan=True)
        # the device-resident entry point reports the same bits
        dev_out = ex.alloc_outputs(5)
        ex.run_device(_gpu.to_device(fovs), _gpu.to_device(np.clip(givens, 0, 2**31 - 1).astype(np.int32)), dev_out)
        dst = _gpu.to_host(dev_out["status"])
        assert dst[1] == st[1] and dst[3] == st[3] and dst[2] == _lib.AMT_FOV_GIVEN_VALUE_RANGE and not dst[0] and not dst[4]
        with pytest.raises(FovCapacityError):
            ex.check_status(dev_out)


def test_host_label_masks_give_the_same_answer_on_every_route():
    """amt_executor_run_host sends host label masks across PCIe as per-row runs of equal value (encoded by host threads
    into pinned staging, decoded on the device; amt_tune('exec_host_rle')); with that off, int64 masks are narrowed to
    uint16 by host threads when max_label_value < 65535 (amt_tune('exec_host_narrow')) or cross as they are and are
    narrowed on the device.  Same tables, counts and status bits on all three routes and for all three mask dtypes,
    including a value beyond uint16, one beyond int32 and a negative one; large enough (2 x 1024 x 1024) for the host
    threads to split a chunk."""
    from arcadia_microscopy_tools_b200 import _lib

    lib = _lib.load()
    C, shape, cells = 2, (1024, 1024), 300
    fovs, givens = [], []
    for i in range(3):
        f, g, _ = make_fov(4100 + i, C, shape[0], shape[1], cells)
        fovs.append(f), givens.append(g.astype(np.int64))
    fovs, givens = np.stack(fovs), np.stack(givens)
    clean = givens.copy()
    givens[1, 500, 500] = 70000
    givens[1, 600, 600] = 2**40
    givens[2, 1023, 1023] = -7
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=0, chunk_fovs=2, max_labels=512,
                            max_label_value=int(givens[0].max()) + 5, given_label_dtype=np.int64)

    def run(config, masks, rle, narrow, share=100):
        _lib.check(lib.amt_tune(b"exec_host_rle", rle))
        _lib.check(lib.amt_tune(b"exec_host_narrow", narrow))
        _lib.check(lib.amt_tune(b"exec_rle_share", share))
        try:
            with FovBatchExecutor(config) as ex:
                out = ex.run_host(fovs, masks, on_error="status")
                assert ex.last_rle_masks == (0 if not rle else (3 if share == 100 else 2))
                return out, ex.last_h2d_bytes, ex.last_plain_mask_chunks
        finally:
            _lib.check(lib.amt_tune(b"exec_host_rle", 1))
            _lib.check(lib.amt_tune(b"exec_host_narrow", 1))
            _lib.check(lib.amt_tune(b"exec_rle_share", -1))

    def same(a, b):
        assert np.array_equal(a["status"], b["status"])
        assert np.array_equal(a["counts_given"], b["counts_given"]) and a["counts_given"].min() > 0
        assert np.array_equal(a["counts_thr"], b["counts_thr"])
        for i in range(3):
            k = int(a["counts_given"][i])
            assert np.array_equal(a["tables_given"][i][:, :k], b["tables_given"][i][:, :k], equal_nan=True)

    (a, bytes_rle, plain), (b, bytes_narrow, _), (c, bytes_i64, _) = run(cfg, givens, 1, 1), run(cfg, givens, 0, 1), run(cfg, givens, 0, 0)
    assert a["status"][0] == 0 and a["status"][1] & _lib.AMT_FOV_GIVEN_VALUE_RANGE and a["status"][2] & _lib.AMT_FOV_GIVEN_NEGATIVE
    same(a, b), same(a, c)
    # half of every chunk's masks as runs, the other half plain behind the images (what the executor does by itself when
    # the host threads are slower than PCIe): FOV 0 and 2 as runs, FOV 1 as int64
    h, bytes_half, _ = run(cfg, givens, 1, 1, share=50)
    same(a, h)
    px = 3 * shape[0] * shape[1]
    assert bytes_half > px * 2 * C + shape[0] * shape[1] * 8
    assert plain == 0 and bytes_narrow == px * (2 * C + 2) and bytes_i64 == px * (2 * C + 8)
    assert px * 2 * C < bytes_rle < px * 2 * C + px // 4  # the masks cross as runs: a small fraction of a byte per pixel
    # the other two host dtypes, run-length staged and plain
    for dtype in (np.int32, np.uint16):
        cfg_d = dataclasses.replace(cfg, given_label_dtype=dtype)
        d, bytes_d, plain_d = run(cfg_d, clean.astype(dtype), 1, 1)
        e, bytes_e, _ = run(cfg_d, clean.astype(dtype), 0, 1)
        same(d, e), same(d, run(cfg_d, clean.astype(dtype), 1, 1, share=50)[0])
        assert plain_d == 0 and px * 2 * C < bytes_d <= bytes_rle and bytes_e == px * (2 * C + np.dtype(dtype).itemsize) and not d["status"].any()
        k = int(a["counts_given"][0])
        assert np.array_equal(d["tables_given"][0][:, :k], a["tables_given"][0][:, :k], equal_nan=True)


def test_ragged_label_masks_fall_back_to_the_plain_route():
    """A label mask with fewer than four pixels per run (here: a different label in every pixel of a block) does not fit
    the run-length staging; its chunk crosses PCIe as the plain mask and gives the same answer, next to chunks that fit."""
    from arcadia_microscopy_tools_b200 import _lib

    lib = _lib.load()
    C, shape = 2, (1024, 1024)
    fovs, givens = [], []
    for i in range(4):
        f, g, _ = make_fov(4200 + i, C, shape[0], shape[1], 200)
        fovs.append(f), givens.append(g.astype(np.int64))
    fovs, givens = np.stack(fovs), np.stack(givens)
    rng = np.random.default_rng(3)
    givens[2] = rng.integers(1, 3, shape) * 400 + np.arange(shape[1]) % 2  # every pixel differs from its neighbour
    givens[2, :4] = 0
    givens[2, 7, 9] = -1
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=0, chunk_fovs=1, max_labels=2048,
                            max_label_value=1000, given_label_dtype=np.int64)
    outs = []
    for rle in (1, 0):
        _lib.check(lib.amt_tune(b"exec_host_rle", rle))
        _lib.check(lib.amt_tune(b"exec_rle_share", 100))  # every mask is tried as runs (no balancing against PCIe here)
        try:
            with FovBatchExecutor(cfg) as ex:
                outs.append(ex.run_host(fovs, givens, on_error="status"))
                assert ex.last_plain_mask_chunks == (1 if rle else 0) and ex.last_rle_masks == (3 if rle else 0)
        finally:
            _lib.check(lib.amt_tune(b"exec_host_rle", 1))
            _lib.check(lib.amt_tune(b"exec_rle_share", -1))
    a, b = outs
    assert np.array_equal(a["status"], b["status"]) and a["status"][2] & _lib.AMT_FOV_GIVEN_NEGATIVE and not a["status"][3]
    assert np.array_equal(a["counts_given"], b["counts_given"])
    for i in range(4):
        k = min(int(a["counts_given"][i]), cfg.max_labels)
        assert np.array_equal(a["tables_given"][i][:, :k], b["tables_given"][i][:, :k], equal_nan=True)


def test_run_host_from_library_pinned_staging():
    """amt_host_alloc staging (plain and write-combined) feeds the host-fed entry point like any host buffer."""
    n_fov, C, shape = 2, 2, (128, 160)
    fovs = np.stack([make_fov(300 + i, C, shape[0], shape[1], 25)[0] for i in range(n_fov)])
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=0, chunk_fovs=1, max_labels=256,
                            quantify_given_mask=False)
    with FovBatchExecutor(cfg) as ex:
        want = ex.run_host(fovs, None)
        for wc in (False, True):
            with _gpu.PinnedBuffer(fovs.shape, np.uint16, write_combined=wc) as staging:
                staging.array[...] = fovs
                got = ex.run_host(staging.array, None)
                assert np.array_equal(got["counts_thr"], want["counts_thr"]) and np.array_equal(got["thresholds"], want["thresholds"])
                for i in range(n_fov):
                    k = int(want["counts_thr"][i])
                    assert np.array_equal(got["tables_thr"][i][:, :k], want["tables_thr"][i][:, :k], equal_nan=True)


def test_decision_exact_retries_planes_with_massive_ties():
    """Decision-exact mode lists the samples within the filter's error bound of a deciding value; a plane with huge
    ties (here: a flat thresholded channel with one bright square, so that thousands of samples share the percentile
    values) overflows those lists, and the executor recomputes that field of view with the float64 kernels.  Results
    equal the float64 mode's bit for bit either way, and the ordinary FOVs next to it are not retried."""
    C, shape = 2, (160, 192)
    fovs, givens = [], []
    for i in range(4):
        f, g, _ = make_fov(5100 + i, C, shape[0], shape[1], 25)
        fovs.append(f), givens.append(g)
    fovs, givens = np.stack(fovs), np.stack(givens)
    fovs[2, 0] = 300
    fovs[2, 0, 0:10, 0:12] = 5000  # far from it the filtered plane is perfectly flat: > 20 000 tied samples
    outs = {}
    for seg_filter in ("decision_exact", "float64"):
        cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=0, chunk_fovs=2, max_labels=256,
                                max_label_value=int(givens.max()), seg_plane_filter=seg_filter)
        with FovBatchExecutor(cfg) as ex:
            dev_out = ex.alloc_outputs(4, labels=True)
            ex.run_device(_gpu.to_device(fovs), _gpu.to_device(givens), dev_out)
            retries_device = ex.retry_count
            host_out = ex.run_host(fovs, givens)
            outs[seg_filter] = ({k: _gpu.to_host(v) for k, v in dev_out.items() if v is not None}, host_out,
                                retries_device, ex.retry_count)
    dx, f64 = outs["decision_exact"], outs["float64"]
    assert dx[2] == 1 and dx[3] == 2 and f64[3] == 0  # FOV 2 retried once per entry point, nobody else
    for key in ("thresholds", "counts_thr", "counts_given", "labels_thr", "labels_given"):
        assert np.array_equal(dx[0][key], f64[0][key]), key
    for which in ("thr", "given"):
        for i in range(4):
            k = int(f64[0][f"counts_{which}"][i])
            assert np.array_equal(dx[0][f"tables_{which}"][i][:, :k], f64[0][f"tables_{which}"][i][:, :k], equal_nan=True)
            assert np.array_equal(dx[1][f"tables_{which}"][i][:, :k], f64[1][f"tables_{which}"][i][:, :k], equal_nan=True)
    assert np.array_equal(dx[1]["thresholds"], f64[1]["thresholds"]) and np.array_equal(dx[1]["counts_thr"], f64[1]["counts_thr"])


def test_nd2_file_to_pinned_staging_to_executor(tmp_path):
    """SURVEY 8f-3, end to end: raw ND2 frames -> pinned staging (one memcpy per frame) -> one H2D copy ->
    de-interleave kernel -> the executor's device-resident entry point; the same tables as the host reader's planar
    array through the host-fed entry point."""
    from nd2_synth import write_nd2

    from arcadia_microscopy_tools_b200 import nd2_raw

    C, shape = 4, (128, 160)
    frames = np.stack([make_fov(6100 + i, C, shape[0], shape[1], 25)[0] for i in range(3)])
    path = tmp_path / "plate.nd2"
    write_nd2(path, frames)
    cfg = FovPipelineConfig(n_channels=C, height=shape[0], width=shape[1], seg_channel=1, chunk_fovs=2, max_labels=256,
                            quantify_given_mask=False)
    with FovBatchExecutor(cfg) as ex:
        dev_frames = nd2_raw.read_nd2_to_device(path)
        assert tuple(dev_frames.shape) == (3, C, *shape)
        out = ex.alloc_outputs(3)
        ex.run_device(dev_frames, None, out)
        got = {k: _gpu.to_host(v) for k, v in out.items() if v is not None}
        want = ex.run_host(nd2_raw.read_nd2_frames(path), None)
    assert np.array_equal(got["counts_thr"], want["counts_thr"]) and got["counts_thr"].min() > 0
    assert np.array_equal(got["thresholds"], want["thresholds"])
    for i in range(3):
        k = int(want["counts_thr"][i])
        assert np.array_equal(got["tables_thr"][i][:, :k], want["tables_thr"][i][:, :k], equal_nan=True)
