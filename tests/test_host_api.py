"""Host-side behaviour of the drop-in API (no GPU needed): the reference's own
``tests/test_pipeline.py`` expectations for ImageOperation / Pipeline, MicroscopyImage
validation and channel slicing, SegmentationMask argument checks, and the C-ABI surface."""

from __future__ import annotations

import ctypes
import re
import warnings
from pathlib import Path

import numpy as np
import pytest

import arcadia_microscopy_tools_b200 as amt
from arcadia_microscopy_tools_b200 import _lib, masks, operations
from arcadia_microscopy_tools_b200.channels import BRIGHTFIELD, DAPI, FITC, TRITC, Channel
from arcadia_microscopy_tools_b200.pipeline import ImageOperation, Pipeline

ROOT = Path(__file__).resolve().parents[1]


def double_intensity(x):
    return x * 2


def add_ten(x):
    return x + 10


def to_float_normalized(x):
    return x.astype(float) / x.max()


# ---------------------------------------------------------------- ImageOperation
def test_image_operation_stores_and_calls():
    op = ImageOperation(np.add, 5)
    assert op.func == np.add and op.args == (5,) and op.kwargs == {}
    assert ImageOperation(np.clip, a_min=0, a_max=100).kwargs == {"a_min": 0, "a_max": 100}
    np.testing.assert_array_equal(op(np.array([1, 2, 3])), [6, 7, 8])
    assert "double_intensity" in repr(ImageOperation(double_intensity))


def test_image_operation_is_immutable_and_hashable():
    op = ImageOperation(double_intensity)
    with pytest.raises(AttributeError):
        op.func = add_ten
    with pytest.raises(AttributeError):
        del op.func
    assert op == ImageOperation(double_intensity) and op != ImageOperation(add_ten)
    assert ImageOperation(np.add, 5) == ImageOperation(np.add, 5) != ImageOperation(np.add, 10)
    assert hash(ImageOperation(np.add, 5, k=1)) == hash(ImageOperation(np.add, 5, k=1))


# ---------------------------------------------------------------- Pipeline
def test_pipeline_defaults_and_validation():
    p = Pipeline(operations=[ImageOperation(double_intensity), ImageOperation(add_ten)])
    assert len(p) == 2 and p.copy is False and p.preserve_dtype is False and p.parallel is False
    assert p.max_workers is None
    with pytest.raises(ValueError, match="at least one operation"):
        Pipeline(operations=[])
    with pytest.raises(ValueError, match="at least one operation"):
        Pipeline(operations=[], parallel=True)
    for bad in (0, -1):
        with pytest.raises(ValueError, match="max_workers must be at least 1"):
            Pipeline(operations=[ImageOperation(double_intensity)], max_workers=bad)
    with pytest.raises(TypeError, match="All operations must be callable"):
        Pipeline(operations=("not_a_function",))
    with pytest.raises(TypeError, match="All operations must be callable"):
        Pipeline(operations=(ImageOperation(double_intensity), 42))
    assert isinstance(Pipeline(operations=(ImageOperation(double_intensity),)).operations, list)
    assert "parallel=True" in repr(Pipeline([ImageOperation(add_ten)], parallel=True))


def test_pipeline_sequential_semantics():
    img = np.array([1, 2, 3], dtype=np.uint16)
    out = Pipeline([ImageOperation(double_intensity), ImageOperation(add_ten)])(img)
    np.testing.assert_array_equal(out, [12, 14, 16])
    assert out.dtype == np.uint16
    f = Pipeline([ImageOperation(to_float_normalized)])(np.array([10, 20, 30], dtype=np.uint16))
    np.testing.assert_allclose(f, [1 / 3, 2 / 3, 1.0])
    keep = Pipeline([ImageOperation(to_float_normalized)], preserve_dtype=True)(np.array([10, 20, 30], dtype=np.uint16))
    assert keep.dtype == np.uint16


def test_pipeline_parallel_semantics():
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        Pipeline([ImageOperation(double_intensity)], parallel=True, copy=True)
        assert len(w) == 1 and "copy=True has no effect" in str(w[0].message)
    p = Pipeline([ImageOperation(double_intensity)], parallel=True)
    for bad in (np.zeros((2, 2), np.uint16), np.zeros(3, np.uint16)):
        with pytest.raises(ValueError, match="at least 3D input"):
            p(bad)
    stack = np.random.default_rng(0).integers(0, 100, size=(10, 32, 32)).astype(np.uint16)
    out = Pipeline([ImageOperation(double_intensity), ImageOperation(add_ten)], parallel=True, max_workers=2)(stack)
    np.testing.assert_array_equal(out, stack * 2 + 10)
    assert out.dtype == np.uint16
    one = np.array([[[10, 20], [30, 40]]], dtype=np.uint16)
    assert Pipeline([ImageOperation(to_float_normalized)], parallel=True)(one).dtype == np.float64
    assert Pipeline([ImageOperation(to_float_normalized)], parallel=True, preserve_dtype=True)(one).dtype == np.uint16


def test_device_ops_are_recognised_without_touching_the_gpu():
    ops = [ImageOperation(operations.subtract_background_dog, low_sigma=1, high_sigma=10),
           ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99))]
    assert all(op.runs_on_device for op in ops) and Pipeline(ops)._device_chain()
    assert not Pipeline([ops[0], ImageOperation(double_intensity)])._device_chain()


# ---------------------------------------------------------------- operations: host-side contract
def test_operation_argument_validation_matches_reference_messages():
    x = np.arange(12, dtype=np.uint16).reshape(3, 4)
    with pytest.raises(ValueError, match="Invalid percentile range"):
        operations.rescale_by_percentile(x, (50, 10))
    with pytest.raises(ValueError, match="Percentile must be between 0 and 100"):
        operations.subtract_background_dog(x, percentile=-1)
    with pytest.raises(ValueError, match=re.escape("low_sigma (3) must be smaller than high_sigma (2)")):
        operations.subtract_background_dog(x, 3, 2)
    empty = np.zeros((0, 4), np.uint16)
    assert operations.rescale_by_percentile(empty).dtype == np.float64
    assert operations.apply_threshold(empty).dtype == np.bool_
    with pytest.raises(ValueError, match="Unsupported thresholding method: 'nope'"):
        operations.apply_threshold(x, method="nope")
    # a constant image is all False before the method name is looked at (operations.py:201-209)
    assert not operations.apply_threshold(np.full((3, 3), 5, np.uint16), method="nope").any()


def test_crop_to_center_is_a_view():
    x = np.arange(4 * 10 * 12).reshape(4, 10, 12)
    c = operations.crop_to_center(x, (4, 6))
    assert c.shape == (4, 4, 6) and c.base is not None
    np.testing.assert_array_equal(c, x[:, 3:7, 3:9])
    assert operations.crop_to_center(x, (100, 100)).shape == x.shape


# ---------------------------------------------------------------- MicroscopyImage
def _image(shape=(4, 8, 8), axes="CYX", channels=(BRIGHTFIELD, DAPI, FITC, TRITC)):
    data = np.arange(int(np.prod(shape)), dtype=np.uint16).reshape(shape)
    return amt.MicroscopyImage.from_arrays(data, list(channels), axes)


def test_microscopy_image_channel_views():
    im = _image()
    assert im.num_channels == 4 and im.channel_axis == 0 and im.dimensions.is_multichannel
    v = im.get_channel_intensities("FITC")
    assert v.shape == (8, 8) and np.shares_memory(v, im.intensities)
    np.testing.assert_array_equal(v, im.intensities[2])
    np.testing.assert_array_equal(im.get_channel_intensities(DAPI), im.intensities[1])
    with pytest.raises(ValueError, match="Channel 'CY5' not found in image. Available channels"):
        im.get_channel_intensities("CY5")
    tc = _image((5, 2, 8, 8), "TCYX", (DAPI, FITC))
    assert tc.channel_axis == 1 and tc.get_channel_intensities(FITC).shape == (5, 8, 8)
    assert not tc.get_channel_intensities(FITC).flags.c_contiguous
    single = _image((3, 8, 8), "TYX", (FITC,))
    assert single.get_channel_intensities(FITC) is single.intensities
    out = im.apply_pipeline(Pipeline([ImageOperation(double_intensity)]), "DAPI")
    np.testing.assert_array_equal(out, im.intensities[1] * 2)


def test_microscopy_image_validation():
    im = _image()
    with pytest.raises(ValueError, match="does not match"):
        amt.MicroscopyImage(np.zeros((4, 8, 9), np.uint16), im.metadata)
    with pytest.warns(amt.MetadataWarning, match="Expected uint16"):
        amt.MicroscopyImage(im.intensities.astype(np.float32), im.metadata)
    with pytest.raises(ValueError, match="hex code"):
        Channel("X", "red")
    assert "MicroscopyImage" in repr(im) and "DAPI" in repr(im.metadata)


# ---------------------------------------------------------------- SegmentationMask: checks before any GPU work
def test_segmentation_mask_validation_messages():
    good = np.zeros((6, 6), np.int64)
    good[2:4, 2:4] = 1
    with pytest.raises(TypeError, match="mask_image must be a numpy array"):
        masks.SegmentationMask([[0, 1]])
    with pytest.raises(ValueError, match="must be a 2D array"):
        masks.SegmentationMask(np.zeros((2, 3, 3), np.int64))
    with pytest.raises(ValueError, match="non-negative"):
        masks.SegmentationMask(good - 1)
    with pytest.raises(ValueError, match="contains no cells"):
        masks.SegmentationMask(np.zeros((4, 4), np.int64))
    with pytest.raises(TypeError, match="must be a Mapping"):
        masks.SegmentationMask(good, intensity_image_dict=[1])
    with pytest.raises(ValueError, match="same shape as mask_image"):
        masks.SegmentationMask(good, intensity_image_dict={DAPI: np.zeros((5, 5), np.uint16)})
    with pytest.raises(ValueError, match="must be 2D"):
        masks.SegmentationMask(good, intensity_image_dict={DAPI: np.zeros((6, 6, 1), np.uint16)})
    m = masks.SegmentationMask(good, intensity_image_dict={DAPI: np.zeros((6, 6), np.uint16)}, remove_edge_cells=False)
    assert m.property_names == masks.DEFAULT_CELL_PROPERTY_NAMES
    assert m.intensity_property_names == masks.DEFAULT_INTENSITY_PROPERTY_NAMES
    assert masks.SegmentationMask(good).intensity_property_names == []
    with pytest.raises(AttributeError, match="Cannot modify 'mask_image'"):
        m.mask_image = good
    with pytest.warns(UserWarning, match="Centroid property not available"):
        empty = masks.SegmentationMask(good, property_names=["label"]).centroids_yx
    assert empty.shape == (0, 2)


# ---------------------------------------------------------------- C ABI surface
def test_shared_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "amt_b200.h").read_text()
    declared = set(re.findall(r"\b(amt_[a-z0-9_]+)\s*\(", header))
    declared -= {"amt_executor", "amt_fov_config", "amt_map_params"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # raises if the .so is missing: there is no CPU fallback
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.amt_version() >= 100
    assert lib.amt_strerror(_lib.AMT_ERR_CAPACITY) == b"capacity exceeded"
    assert ctypes.sizeof(_lib.MapParams) == 64 and ctypes.sizeof(_lib.FovConfig) == 112


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = np.arange(64, dtype=np.uint16).reshape(8, 8)
    with pytest.raises(_lib.AmtLibraryError, match="no CPU fallback"):
        operations.rescale_by_percentile(x)
    m = np.zeros((8, 8), bool)
    m[2:5, 2:5] = True
    with pytest.raises(_lib.AmtLibraryError, match="no CPU fallback"):
        masks.SegmentationMask(m).label_image


def test_nd2_raw_reader_and_frame_layout(tmp_path):
    """Host side of the ND2 fast path: chunk map, attributes, zero-copy frame offsets."""
    from nd2_synth import write_nd2

    from arcadia_microscopy_tools_b200 import nd2_raw

    rng = np.random.default_rng(3)
    frames = rng.integers(0, 65535, size=(3, 4, 20, 28)).astype(np.uint16)
    path = tmp_path / "synthetic.nd2"
    write_nd2(path, frames)
    assert np.array_equal(nd2_raw.read_nd2_frames(path), frames)
    mm, offsets, (h, w, c) = nd2_raw.nd2_frame_layout(path)
    assert (h, w, c) == (20, 28, 4) and len(offsets) == 3
    for i, off in enumerate(offsets):
        raw = np.frombuffer(mm[off : off + h * w * c * 2].tobytes(), dtype="<u2").reshape(h, w, c)
        assert np.array_equal(raw.transpose(2, 0, 1), frames[i])
    ref_dir = Path("/root/reference/src/arcadia_microscopy_tools/tests/data")
    for f in sorted(ref_dir.glob("*.nd2")) if ref_dir.exists() else []:
        want = nd2_raw.read_nd2_frames(f)
        mm, offsets, (h, w, c) = nd2_raw.nd2_frame_layout(f)
        got = np.stack([np.frombuffer(mm[o : o + h * w * c * 2].tobytes(), dtype="<u2").reshape(h, w, c).transpose(2, 0, 1)
                        for o in offsets])
        assert np.array_equal(got, want), f.name


# ---------------------------------------------------------------- threshold scans (host plan step)
def _bimodal(rng, shape, scale=1.0):
    fg = rng.random(shape) < 0.3
    return np.where(fg, rng.normal(900 * scale, 70 * scale, shape), rng.gamma(2.0, 70.0 * scale, shape)).clip(0, 65535)


def test_histogram_scans_of_the_product_match_the_oracle_restatements():
    """The product's host scans over a device histogram (li, minimum, triangle, otsu twin) against the
    oracle's restatements of scikit-image, on exact integer histograms and 256-bin float histograms."""
    from oracle import threshold as T

    rng = np.random.default_rng(2026)
    for trial in range(5):
        base = _bimodal(rng, (150, 130), 1.0 + 0.3 * trial)
        u16 = base.astype(np.uint16)
        counts, centers = T.histogram_int(u16)
        assert operations._triangle_from_histogram(counts, centers) == T.threshold_triangle(u16)
        assert operations._minimum_from_histogram(counts, centers) == T.threshold_minimum(u16)
        assert operations._li_from_histogram(counts, centers) == T.threshold_li(u16.copy())
        assert operations._li_from_histogram(counts, centers, initial_guess=float(u16.mean()) * 1.5, tolerance=0.1) == \
            T.threshold_li(u16.copy(), initial_guess=float(u16.mean()) * 1.5, tolerance=0.1)
        assert operations._otsu_from_histogram(counts, centers) == T.threshold_otsu(u16)
        f64 = base / 65535.0
        counts, centers = T.histogram_float_256(f64)
        assert operations._triangle_from_histogram(counts, centers) == T.threshold_triangle(f64)
        assert operations._minimum_from_histogram(counts, centers) == T.threshold_minimum(f64)
    flat = rng.poisson(20.0, (64, 64)).astype(np.uint16)  # one mode: no two maxima survive the smoothing
    counts, centers = T.histogram_int(flat)
    with pytest.raises(RuntimeError, match="Unable to find two maxima"):
        T.threshold_minimum(flat)
    with pytest.raises(RuntimeError, match="Unable to find two maxima"):
        operations._minimum_from_histogram(counts, centers)


def test_three_bin_mean_is_scipys_uniform_filter_bit_for_bit():
    from scipy import ndimage as ndi

    rng = np.random.default_rng(8)
    for n in (2, 3, 7, 256, 65536):
        line = rng.integers(0, 90000, n).astype(np.float32)
        for _ in range(6):
            want = ndi.uniform_filter1d(line, 3)
            assert np.array_equal(operations._uniform3_float32(line), want), n
            line = want


def test_threshold_keyword_arguments_are_checked():
    with pytest.raises(TypeError, match="unexpected keyword argument 'window_size'"):
        operations._check_threshold_kwargs("otsu", {"window_size": 5})
    operations._check_threshold_kwargs("yen", {"nbins": 64})
    with pytest.raises(ValueError, match="nbins"):
        operations._check_threshold_kwargs("yen", {"nbins": 1})
    operations._check_threshold_kwargs("li", {"tolerance": 0.25, "initial_guess": 300.0})
    wide = np.array([[-70000, 3], [5, 9]], dtype=np.int64)
    with pytest.raises(NotImplementedError, match="65536 bins"):
        operations._shift_wide_integers(wide)
    shifted, offset = operations._shift_wide_integers(np.array([[-5, 3], [5, 9]], dtype=np.int32))
    assert shifted.dtype == np.uint16 and offset == -5 and shifted.tolist() == [[0, 8], [10, 14]]


# ---------------------------------------------------------------- LIF raw reader (host I/O)
def test_lif_raw_reader_round_trip(tmp_path):
    """ref: leica.py:39-80 (list_image_names, load_lif_image -> liffile ... asarray()): the pixel block of a named
    image in the axis order the file stores, for both container versions."""
    from lif_synth import write_lif

    from arcadia_microscopy_tools_b200 import lif_raw
    from arcadia_microscopy_tools_b200.microscopy import MicroscopyImage

    rng = np.random.default_rng(4)
    zstack = rng.integers(0, 65535, (5, 3, 16, 24), dtype=np.uint16)      # Z, C, Y, X (Stellaris-shaped)
    plane8 = rng.integers(0, 255, (12, 10), dtype=np.uint8)               # Y, X
    lapse = rng.integers(0, 4095, (4, 2, 8, 8), dtype=np.uint16)          # T, C, Y, X
    for version in (1, 2):
        path = tmp_path / f"synthetic_v{version}.lif"
        write_lif(path, [("zstack", zstack, "ZCYX"), ("overview", plane8, "YX"), ("lapse", lapse, "TCYX")], version=version)
        assert lif_raw.list_image_names(path) == ["zstack", "overview", "lapse"]
        got, sizes = lif_raw.read_lif_image(path, "zstack")
        assert sizes == {"Z": 5, "C": 3, "Y": 16, "X": 24} and got.dtype == np.uint16 and np.array_equal(got, zstack)
        got, sizes = lif_raw.read_lif_image(path, "overview")
        assert sizes == {"Y": 12, "X": 10} and got.dtype == np.uint8 and np.array_equal(got, plane8)
        info = lif_raw.lif_image_info(path, "lapse")
        assert list(info.sizes) == ["T", "C", "Y", "X"] and info.memory_size == lapse.nbytes
        with pytest.raises(ValueError, match="Image missing not found in .* Available images"):
            lif_raw.read_lif_image(path, "missing")
        image = MicroscopyImage.from_lif_path(path, "zstack", channels=[DAPI, FITC, TRITC])
        assert image.sizes == {"Z": 5, "C": 3, "Y": 16, "X": 24}
        assert np.array_equal(image.get_channel_intensities(FITC), zstack[:, 1])
        with pytest.raises(ValueError, match="channels must be given"):
            MicroscopyImage.from_lif_path(path, "lapse")
        single = MicroscopyImage.from_lif_path(path, "overview")
        assert single.sizes == {"Y": 12, "X": 10} and single.intensities.dtype == np.uint8
    # channel-interleaved storage (channel stride = one sample): the axes come back in stored order, C last
    interleaved = rng.integers(0, 60000, (6, 7, 2), dtype=np.uint16)
    inter_path = tmp_path / "interleaved.lif"
    write_lif(inter_path, [("inter", interleaved, "YXC")])
    got, sizes = lif_raw.read_lif_image(inter_path, "inter")
    assert sizes == {"Y": 6, "X": 7, "C": 2} and np.array_equal(got, interleaved)
    image = MicroscopyImage.from_lif_path(inter_path, "inter", channels=[DAPI, FITC])
    assert np.array_equal(image.get_channel_intensities(FITC), interleaved[..., 1])
    nested = tmp_path / "nested.lif"
    write_lif(nested, [("a", plane8, "YX")], folder="Project")
    assert lif_raw.list_image_names(nested) == ["Project/a"]
    assert np.array_equal(lif_raw.read_lif_image(nested, "Project/a")[0], plane8)
    broken = tmp_path / "broken.lif"
    broken.write_bytes(b"\\x00" * 64)
    with pytest.raises(ValueError, match="not a LIF file"):
        lif_raw.list_image_names(broken)
    truncated = tmp_path / "truncated.lif"
    truncated.write_bytes((tmp_path / "synthetic_v2.lif").read_bytes()[:-100])
    with pytest.raises(ValueError):
        lif_raw.read_lif_image(truncated, "lapse")


def test_histogram_scans_against_independent_constructions():
    """The host scans of the product checked against something that is NOT their twin in oracle/: OpenCV's Otsu and
    Triangle on uint8 images, and brute-force evaluations of the published criteria on small histograms
    (isodata: the threshold t with t == (mean_below + mean_above) / 2; yen: argmax of the entropic criterion written
    out term by term; minimum: the valley between the two surviving peaks of a hand-built bimodal histogram)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for trial in range(6):
        img = np.clip(np.where(rng.random((96, 96)) < 0.35, rng.normal(170, 18, (96, 96)), rng.normal(60, 25 + 3 * trial, (96, 96))), 0, 255).astype(np.uint8)
        lo, hi = int(img.min()), int(img.max())
        counts = np.bincount(img.ravel(), minlength=256)[lo : hi + 1]
        centers = np.arange(lo, hi + 1)
        t_cv, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert np.array_equal(img > operations._otsu_from_histogram(counts, centers), img > t_cv)
        t_tri, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_TRIANGLE)
        assert abs(float(operations._triangle_from_histogram(counts, centers)) - t_tri) <= 1  # OpenCV's end convention differs by one level
        # isodata, brute force over every candidate threshold: the lowest t with t >= (m_below + m_above) / 2 midpoint rule
        c = counts.astype(np.float64)
        best = None
        for k in range(len(centers) - 1):
            w_lo, w_hi = c[: k + 1].sum(), c[k + 1 :].sum()
            if w_lo == 0 or w_hi == 0:
                continue
            m_lo = (c[: k + 1] * centers[: k + 1]).sum() / w_lo
            m_hi = (c[k + 1 :] * centers[k + 1 :]).sum() / w_hi
            if centers[k] >= (m_lo + m_hi) / 2.0 - 1.0:  # bin width 1: the first centre within one bin of the midpoint
                best = centers[k]
                break
        got = operations._isodata_from_histogram(counts, centers)
        assert best is not None and abs(float(got) - float(best)) <= 1.0
        # yen, term by term
        pmf = c / c.sum()
        crit = []
        for k in range(len(centers) - 1):
            p1 = pmf[: k + 1].sum()
            crit.append(-np.log((pmf[: k + 1] ** 2).sum() * (pmf[k + 1 :] ** 2).sum()) + 2 * np.log(p1 * (1 - p1)))
        assert operations._yen_from_histogram(counts, centers) == centers[int(np.argmax(crit))]
    # minimum: two clean triangular peaks at 20 and 70 with a flat valley bottom at 40..45 -> the first lowest bin
    hist = np.zeros(100, dtype=np.int64)
    for k in range(100):
        hist[k] = max(0, 50 - 5 * abs(k - 20)) + max(0, 80 - 4 * abs(k - 70))
    t = operations._minimum_from_histogram(hist, np.arange(100))
    assert 30 <= t <= 50 and hist[int(t)] == hist[30:51].min()


def test_nd2_reader_honours_row_pitch_and_rejects_bad_offsets(tmp_path):
    """ND2 rows are padded to four bytes (uiWidthBytes): an odd width x components must not shear the image; chunk
    offsets and lengths read from the file are bounds-checked before use."""
    from nd2_synth import write_nd2

    from arcadia_microscopy_tools_b200 import nd2_raw

    rng = np.random.default_rng(3)
    frames = rng.integers(0, 65536, size=(2, 3, 5, 7)).astype(np.uint16)  # 7 * 3 * 2 = 42 bytes per row -> pitch 44
    path = tmp_path / "padded.nd2"
    write_nd2(path, frames, row_align=4)
    assert np.array_equal(nd2_raw.read_nd2_frames(path), frames)
    with pytest.raises(ValueError, match="padded"):
        nd2_raw.nd2_frame_layout(path)
    tight = tmp_path / "tight.nd2"
    write_nd2(tight, frames)
    assert np.array_equal(nd2_raw.read_nd2_frames(tight), frames)
    blob = bytearray(tight.read_bytes())
    bad = tmp_path / "bad.nd2"
    bad.write_bytes(bytes(blob[:-8]) + (10**12).to_bytes(8, "little"))  # chunk-map offset far beyond the file
    with pytest.raises(ValueError, match="outside the file"):
        nd2_raw.read_nd2_frames(bad)
    bad.write_bytes(bytes(blob[: len(blob) // 2]))  # truncated
    with pytest.raises(ValueError):
        nd2_raw.read_nd2_frames(bad)
