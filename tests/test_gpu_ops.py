"""GPU parity: operations.* and Pipeline against the CPU oracle (bit-exact for float64
planes and masks: every arithmetic step is the separately rounded operation NumPy/SciPy do)."""

from __future__ import annotations

import hashlib

import numpy as np
import pytest

import oracle
from oracle import filters

pytestmark = pytest.mark.gpu

from arcadia_microscopy_tools_b200 import _gpu, operations  # noqa: E402
from arcadia_microscopy_tools_b200.pipeline import ImageOperation, Pipeline  # noqa: E402
from arcadia_microscopy_tools_b200.synthetic import make_fov  # noqa: E402


def _bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if not np.array_equal(a, b):
        bad = np.flatnonzero(a.ravel() != b.ravel())
        diff = np.abs(a.ravel()[bad].astype(np.float64) - b.ravel()[bad].astype(np.float64))
        raise AssertionError(f"{what}: {bad.size}/{a.size} differ, max |d|={diff.max():.3e}, first idx {bad[:5]}")


@pytest.mark.parametrize("sigma", [0.6, 1.0, 2.7, 16.0])
@pytest.mark.parametrize("shape", [(64, 64), (97, 131), (300, 70)])
def test_gaussian_axis_passes_bit_exact(sigma, shape):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 65535, size=shape).astype(np.uint16)
    dev = _gpu.to_device(img)
    got = _gpu.to_host(_gpu.gaussian_nd(dev, 1.0 / 65535.0, sigma))
    _bits_equal(got, filters.gaussian(img, sigma), f"gaussian sigma={sigma}")
    f = rng.normal(size=shape)
    got = _gpu.to_host(_gpu.gaussian_nd(_gpu.to_device(f), 1.0, sigma))
    _bits_equal(got, filters.gaussian(f, sigma), f"gaussian f64 sigma={sigma}")


def test_gaussian_3d_and_1d():
    rng = np.random.default_rng(12)
    vol = rng.integers(0, 4000, size=(7, 45, 38)).astype(np.uint16)
    got = _gpu.to_host(_gpu.gaussian_nd(_gpu.to_device(vol), 1.0 / 65535.0, 2.0))
    _bits_equal(got, filters.gaussian(vol, 2.0), "3-D gaussian")
    line = rng.random(1000)
    got = _gpu.to_host(_gpu.gaussian_nd(_gpu.to_device(line), 1.0, 5.0))
    _bits_equal(got, filters.gaussian(line, 5.0), "1-D gaussian")


@pytest.mark.parametrize("shape", [(256, 256), (100, 333), (17, 40)])
@pytest.mark.parametrize("sigmas", [(0.6, 16.0), (1.0, 10.0), (2.0, 3.5)])
def test_dog2d_bit_exact(shape, sigmas):
    rng = np.random.default_rng(13)
    stack = rng.integers(100, 5000, size=(3, *shape)).astype(np.uint16)
    dog, mm = _gpu.dog2d(_gpu.to_device(stack), 1.0 / 65535.0, *sigmas)
    got = _gpu.to_host(dog)
    mnmx = _gpu.minmax_values(mm, True)
    for i in range(3):
        want = filters.difference_of_gaussians(stack[i], *sigmas)
        _bits_equal(got[i], want, f"dog plane {i} {shape} {sigmas}")
        assert mnmx[i, 0] == want.min() and mnmx[i, 1] == want.max()


@pytest.mark.parametrize("dtype", [np.uint16, np.float64])
def test_percentiles_exact(dtype):
    rng = np.random.default_rng(14)
    data = (rng.gamma(2.0, 400.0, size=(3, 211, 173))).astype(dtype)
    if dtype == np.float64:
        data[1] = np.clip(data[1] - 900, 0, None)  # >50 % exact zeros (ties at the minimum)
        data[2, :50] = data[2].max()  # ties at the maximum
    planes = _gpu.to_device(data).reshape(3, -1)
    qs = [0, 0.5, 1, 50, 90, 99, 100]
    for lo in range(0, len(qs), 3):
        sub = qs[lo : lo + 3]
        got = _gpu.percentiles(planes, sub)
        for i in range(3):
            want = np.percentile(data[i], sub)
            assert np.array_equal(got[i], want), (dtype, i, sub, got[i], want)


def test_percentiles_hard_distributions():
    rng = np.random.default_rng(15)
    n = 300 * 300
    cases = {
        "outlier": np.concatenate([rng.random(n - 1), [1e300]]),
        "constant": np.full(n, 3.25),
        "two_values": rng.integers(0, 2, n).astype(np.float64),
        "negative": -np.abs(rng.normal(size=n)) * 1e-9,
        "peaked": np.concatenate([rng.normal(0, 1e-9, n - 100), rng.normal(0, 1.0, 100)]),
    }
    for name, arr in cases.items():
        planes = _gpu.to_device(arr.reshape(1, -1))
        got = _gpu.percentiles(planes, [1, 37.5, 99])
        want = np.percentile(arr, [1, 37.5, 99])
        assert np.array_equal(got[0], want), (name, got[0], want)


@pytest.mark.parametrize("shape", [(256, 256), (90, 130)])
def test_subtract_background_dog_bit_exact(shape):
    fov, _, _ = make_fov(21, 2, shape[0], shape[1], 25)
    for pct in (0, 35.5, 90):
        got = operations.subtract_background_dog(fov[1], percentile=pct)
        _bits_equal(got, oracle.subtract_background_dog(fov[1], percentile=pct), f"dog pct={pct}")
    f = fov[0].astype(np.float64) * 0.37
    _bits_equal(operations.subtract_background_dog(f, 1.0, 4.0), oracle.subtract_background_dog(f, 1.0, 4.0), "float input")
    vol = fov[:2]  # 3-D input: every axis is filtered, one global percentile
    _bits_equal(operations.subtract_background_dog(vol, 0.6, 2.0, 10), oracle.subtract_background_dog(vol, 0.6, 2.0, 10), "3-D")


def test_rescale_by_percentile_bit_exact():
    fov, _, _ = make_fov(22, 2, 128, 160, 12)
    for rng_, out in (((0, 100), (0, 1)), ((1, 99), (0, 1)), ((2, 98), (0, 65535)), ((5, 60), (-1, 2))):
        _bits_equal(operations.rescale_by_percentile(fov[0], rng_, out), oracle.rescale_by_percentile(fov[0], rng_, out),
                    f"u16 {rng_} {out}")
    x = oracle.subtract_background_dog(fov[1], percentile=90)  # mostly zeros
    _bits_equal(operations.rescale_by_percentile(x, (1, 99)), oracle.rescale_by_percentile(x, (1, 99)), "clipped f64")
    const = np.full((20, 30), 7, np.uint16)
    _bits_equal(operations.rescale_by_percentile(const, (1, 99), (0.25, 1)), oracle.rescale_by_percentile(const, (1, 99), (0.25, 1)), "constant")
    u8 = (fov[0] >> 4).astype(np.uint8)
    _bits_equal(operations.rescale_by_percentile(u8, (1, 99)), oracle.rescale_by_percentile(u8, (1, 99)), "uint8")


def test_rescale_division_rare_paths_bit_exact():
    """The map kernel divides by the plane constant p2 - p1 with a three-instruction sequence and hands the samples
    it cannot serve (tiny non-zero numerators, a tiny or huge divisor) to __ddiv_rn: planes that hit those paths."""
    rng = np.random.default_rng(5)
    n = 96 * 128
    big = rng.random(n)
    tiny_mix = np.where(rng.random(n) < 0.6, 0.0, big)              # p1 = 0: numerators are the samples themselves
    tiny_mix[rng.choice(n, 200, replace=False)] = rng.random(200) * 1e-200   # far below 2^-400
    tiny_mix[rng.choice(n, 50, replace=False)] = 5e-324                     # the smallest subnormal
    planes = {
        "tiny numerators": tiny_mix,
        "subnormal divisor": rng.random(n) * 1e-310,
        "huge divisor": rng.random(n) * 1e305,
        "ordinary": rng.normal(0.0, 1.0, n),
    }
    for name, flat in planes.items():
        x = flat.reshape(96, 128)
        for pr, out in (((1, 99), (0, 1)), ((0, 100), (0, 1)), ((5, 95), (-2.0, 7.5))):
            _bits_equal(operations.rescale_by_percentile(x, pr, out), oracle.rescale_by_percentile(x, pr, out), f"{name} {pr} {out}")


def test_apply_threshold_matches_oracle():
    fov, _, _ = make_fov(23, 2, 200, 240, 30)
    for c in range(2):
        got = operations.apply_threshold(fov[c])
        assert got.dtype == np.bool_
        _bits_equal(got, oracle.apply_threshold(fov[c]), f"otsu u16 channel {c}")
    P = oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[1]), (1, 99))
    _bits_equal(operations.apply_threshold(P), oracle.apply_threshold(P), "otsu f64")
    assert not operations.apply_threshold(np.full((9, 9), 3.5)).any()
    rng = np.random.default_rng(16)
    noisy = rng.normal(size=(150, 150)) * 1e-3 + 5
    _bits_equal(operations.apply_threshold(noisy), oracle.apply_threshold(noisy), "otsu f64 noise")


def test_golden_config1_preprocess(golden):
    fov = golden["fov"]
    for bg in (0, 90):
        for c in range(4):
            x = operations.subtract_background_dog(fov[c], 0.6, 16.0, percentile=bg)
            P = operations.rescale_by_percentile(x, (1, 99), (0, 1))
            sha = hashlib.sha256(np.ascontiguousarray(P).tobytes()).hexdigest()
            assert sha == str(golden[f"bg{bg}/pre_sha256"][c]), (bg, c)
    raw_t = [int((~operations.apply_threshold(fov[c])).sum()) for c in range(4)]
    assert [65536 - v for v in raw_t] == golden["raw_fg"].tolist()


def test_pipeline_device_chain_equals_op_by_op():
    fov, _, _ = make_fov(24, 3, 96, 128, 10)
    ops = [ImageOperation(operations.subtract_background_dog, low_sigma=1, high_sigma=10),
           ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99), out_range=(0, 1))]
    want = np.stack([oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[i], 1, 10), (1, 99)) for i in range(3)])
    got = Pipeline(ops, parallel=True)(fov)
    _bits_equal(got, want, "parallel device pipeline")
    one = Pipeline(ops)(fov[0])
    _bits_equal(one, want[0], "sequential device pipeline")
    seg = Pipeline(ops + [ImageOperation(operations.apply_threshold)], parallel=True)(fov)
    assert seg.dtype == np.bool_
    _bits_equal(seg, np.stack([oracle.apply_threshold(w) for w in want]), "pipeline + threshold")
    u = Pipeline([ops[1]], parallel=True, preserve_dtype=True)(fov)
    assert u.dtype == np.uint16
    # reference integration tests (test_pipeline.py:264-328): loose range/dtype/shape checks
    img = np.random.default_rng(0).integers(0, 65535, size=(3, 128, 128)).astype(np.uint16)
    r = Pipeline([ImageOperation(operations.rescale_by_percentile, percentile_range=(2, 98), out_range=(0, 1))], parallel=True)(img)
    assert r.dtype == np.float64 and r.min() >= 0 and r.max() <= 1 and r.shape == img.shape


def test_plane_constant_division_matches_ddiv_bitwise():
    """The map kernel divides by a per-plane constant with reciprocal + two FMA corrections
    (map.cu div_const); it must equal IEEE division (what NumPy does) for every input."""
    import ctypes as C

    import torch

    from arcadia_microscopy_tools_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(99)
    n = 1 << 22
    parts_a, parts_b = [], []
    # rescale-like operands: 0 <= a <= b
    b = rng.random(n) * 10.0 ** rng.integers(-6, 3, n)
    parts_a.append(b * rng.random(n)); parts_b.append(b)
    # arbitrary magnitudes and signs inside and outside the fast window (tiny / huge -> __ddiv_rn path)
    e = rng.integers(-1000, 1000, n)
    parts_a.append(np.ldexp(rng.random(n) + 0.5, e) * rng.choice([-1.0, 1.0], n))
    parts_b.append(np.ldexp(rng.random(n) + 0.5, rng.integers(-400, 400, n)))
    # divisors with all-ones / all-zeros mantissas, numerators next to multiples of the divisor
    ones = np.nextafter(np.ldexp(1.0, rng.integers(-20, 20, n)), 0.0)
    k = rng.integers(1, 1 << 20, n).astype(np.float64)
    near = ones * k
    near = np.where(rng.random(n) < 0.5, np.nextafter(near, np.inf), np.nextafter(near, -np.inf))
    parts_a.append(near); parts_b.append(ones)
    parts_a.append(np.array([0.0, -0.0, 1.0, 5e-324, 2.0 ** -401, 2.0 ** -399, 1e308])); parts_b.append(np.full(7, 3.0))
    a = np.concatenate(parts_a); b = np.concatenate(parts_b)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    L.check(lib.amt_selftest_div(_gpu.ptr(da), _gpu.ptr(db), a.size, _gpu.ptr(bad), _gpu.stream_ptr()))
    assert int(bad.item()) == 0, f"{int(bad.item())} of {a.size} quotients differ from IEEE division"


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 96), (160, 64), (256, 256), (100, 332), (150, 68), (34, 36), (66, 1028)])
def test_dog2d_strip_variants_bit_exact(variant, shape):
    """Shapes with width % 4 == 0 and height % 2 == 0 take the strip kernels of dog.cu (narrow last
    strip, short last step); every tuning variant must give scipy's bits, for uint16 and float64
    input, plus the plane min/max."""
    from arcadia_microscopy_tools_b200 import _lib as L

    lib = L.load()
    assert lib.amt_tune(b"dog_variant", variant) == 0
    try:
        rng = np.random.default_rng(31 + variant)
        stack = rng.integers(0, 65535, size=(5, *shape)).astype(np.uint16)
        for sigmas in [(0.6, 16.0), (1.0, 8.0), (0.0, 4.0)]:
            dog, mm = _gpu.dog2d(_gpu.to_device(stack), 1.0 / 65535.0, *sigmas)
            got = _gpu.to_host(dog)
            for i in range(stack.shape[0]):
                want = filters.difference_of_gaussians(stack[i], *sigmas)
                _bits_equal(got[i], want, f"variant {variant} dog {sigmas} plane {i}")
            mnmx = _gpu.minmax_values(mm, True)
            assert np.array_equal(mnmx[:, 0], got.reshape(5, -1).min(1)) and np.array_equal(mnmx[:, 1], got.reshape(5, -1).max(1))
        f = rng.normal(size=(2, *shape))
        dog, _ = _gpu.dog2d(_gpu.to_device(f), 1.0, 0.6, 16.0)
        for i in range(2):
            _bits_equal(_gpu.to_host(dog)[i], filters.difference_of_gaussians(f[i], 0.6, 16.0), f"variant {variant} f64 plane {i}")
    finally:
        lib.amt_tune(b"dog_variant", 1)


@pytest.mark.parametrize("size", [1, 2, 3, 8, 15, 50, 101])
@pytest.mark.parametrize("dtype", [np.uint16, np.float64])
def test_white_tophat_matches_scipy(size, dtype):
    rng = np.random.default_rng(size)
    for shape in [(64, 64), (97, 131), (300, 40), (20, 17)]:
        img = rng.integers(0, 65535, size=shape).astype(dtype)
        if dtype == np.float64:
            img = img / 65535.0 - 0.3
        got = operations.subtract_background_tophat(img, size)
        want = filters.white_tophat(img, size)
        assert got.dtype == want.dtype
        _bits_equal(got, want, f"white_tophat size={size} {shape} {dtype}")


def test_white_tophat_batched_and_3d():
    rng = np.random.default_rng(77)
    stack = rng.integers(0, 4000, size=(4, 90, 70)).astype(np.uint16)
    got = operations.subtract_background_tophat(stack, 9, _batched=True)
    for i in range(4):
        _bits_equal(got[i], filters.white_tophat(stack[i], 9), f"batched plane {i}")
    got3 = operations.subtract_background_tophat(stack, 5)  # one 3-D array: the box spans all three axes
    _bits_equal(got3, filters.white_tophat(stack, 5), "3-D top-hat")


def test_gaussian_smooth_matches_scipy():
    rng = np.random.default_rng(78)
    img = rng.integers(0, 65535, size=(128, 96)).astype(np.uint16)
    _bits_equal(operations.gaussian_smooth(img, 2.0), filters.gaussian(img, 2.0), "gaussian_smooth u16")
    f = rng.normal(size=(3, 40, 50))
    got = operations.gaussian_smooth(f, 1.5, _batched=True)
    for i in range(3):
        _bits_equal(got[i], filters.gaussian(f[i], 1.5), f"gaussian_smooth batched {i}")


def test_nd2_fast_path_matches_host_reader(tmp_path):
    """Raw ND2 frame payloads -> pinned staging -> H2D -> amt_deinterleave_u16 == host de-interleave."""
    from nd2_synth import write_nd2

    from arcadia_microscopy_tools_b200 import nd2_raw

    rng = np.random.default_rng(4)
    for shape in [(2, 4, 256, 256), (3, 1, 64, 80), (1, 3, 33, 47), (1, 2, 1030, 517)]:
        frames = rng.integers(0, 65535, size=shape).astype(np.uint16)
        path = tmp_path / f"s{shape[1]}_{shape[2]}.nd2"
        write_nd2(path, frames)
        dev = nd2_raw.read_nd2_to_device(path)
        got = _gpu.to_host(dev).view(np.uint16)
        assert got.shape == frames.shape
        _bits_equal(got, frames, f"nd2 fast path {shape}")


@pytest.mark.parametrize("method", ["isodata", "yen", "mean"])
def test_histogram_threshold_methods_match_oracle(method):
    rng = np.random.default_rng(55)
    base = (rng.gamma(2.0, 500.0, size=(3, 140, 120)) + 800 * (rng.random((3, 140, 120)) < 0.1) * rng.random((3, 140, 120)) * 8).clip(0, 65535)
    u16 = base.astype(np.uint16)
    for i in range(3):
        got = operations.apply_threshold(u16[i], method)
        assert got.dtype == np.bool_ and np.array_equal(got, oracle.apply_threshold(u16[i], method)), (method, i)
    got = operations.apply_threshold(u16, method, _batched=True)
    for i in range(3):
        assert np.array_equal(got[i], oracle.apply_threshold(u16[i], method))
    if method != "mean":  # mean of float pixels: NumPy's pairwise summation is not reproduced on the device
        f = base / 65535.0
        for i in range(3):
            assert np.array_equal(operations.apply_threshold(f[i], method), oracle.apply_threshold(f[i], method)), (method, i)
    const = np.full((32, 32), 7, dtype=np.uint16)
    assert not operations.apply_threshold(const, method).any()


@pytest.mark.parametrize("method", ["li", "minimum", "triangle"])
def test_li_minimum_triangle_match_oracle(method):
    """the remaining histogram-based methods of the reference's ``apply_threshold``; device
    histogram + comparison, skimage's scalar scan on the host; masks must equal the oracle's bit for bit."""
    rng = np.random.default_rng(77)
    shape = (3, 150, 130)
    fg = rng.random(shape) < 0.3
    base = np.where(fg, rng.normal(900, 70, shape), rng.gamma(2.0, 70.0, shape)).clip(0, 65535)
    u16 = base.astype(np.uint16)
    for i in range(3):
        got = operations.apply_threshold(u16[i], method)
        want = oracle.apply_threshold(u16[i].copy(), method)
        assert got.dtype == np.bool_ and 0 < want.sum() < want.size and np.array_equal(got, want), (method, i)
    got = operations.apply_threshold(u16, method, _batched=True)
    for i in range(3):
        assert np.array_equal(got[i], oracle.apply_threshold(u16[i].copy(), method))
    if method == "li":
        kw = dict(tolerance=0.1, initial_guess=float(u16[0].mean()) * 1.4)
        assert np.array_equal(operations.apply_threshold(u16[0], "li", **kw), oracle.apply_threshold(u16[0].copy(), "li", **kw))
        # float images: scikit-image iterates over the pixels; shapes off every block / tile boundary
        f = base / 65535.0
        for img in (f[0], f[1, :97, :61], f[2, :33, :64] - 0.2, np.round(f[0], 2)):
            want = oracle.apply_threshold(img.copy(), "li")
            assert 0 < want.sum() < want.size and np.array_equal(operations.apply_threshold(img, "li"), want), img.shape
        kw = dict(tolerance=1e-4, initial_guess=float(f[0].mean()) * 1.3)
        assert np.array_equal(operations.apply_threshold(f[0], "li", **kw), oracle.apply_threshold(f[0].copy(), "li", **kw))
        chain = operations.rescale_by_percentile(u16[0], (1, 99))  # the realistic chain the verdict names
        assert np.array_equal(operations.apply_threshold(chain, "li"), oracle.apply_threshold(chain.copy(), "li"))
        with pytest.raises(ValueError, match="initial guess"):
            operations.apply_threshold(f[0], "li", initial_guess=5.0)
        # 'mean' on float64 planes: NumPy's pairwise sum reproduced bit for bit on the device
        for shape2 in ((150, 130), (97, 257), (1, 5), (640, 1000)):
            f = rng.random(shape2) * 3.0 - 1.0
            assert np.array_equal(operations.apply_threshold(f, "mean"), f > np.mean(f)), shape2
            assert np.array_equal(_gpu.plane_sums_f64(_gpu.to_device(f.reshape(1, -1))), np.array([np.sum(f)])), shape2
        chain = operations.apply_threshold(operations.rescale_by_percentile(u16[0], (1, 99)), "mean")
        assert np.array_equal(chain, oracle.apply_threshold(oracle.rescale_by_percentile(u16[0], (1, 99)), "mean"))
    else:
        f = base / 65535.0
        for i in range(3):
            assert np.array_equal(operations.apply_threshold(f[i], method), oracle.apply_threshold(f[i], method)), (method, i)
        # nbins (operations.py:214 forwards it): the device histogram for any bin count equals np.histogram's
        for nbins in (16, 100, 1000):
            (counts, centers), = _gpu.plane_histograms(_gpu.to_device(f[0].reshape(1, -1)), nbins=nbins)
            want_counts, edges = np.histogram(f[0].ravel(), bins=nbins)
            assert np.array_equal(counts, want_counts) and np.array_equal(centers, (edges[:-1] + edges[1:]) / 2)
            assert np.array_equal(operations.apply_threshold(f[0], method, nbins=nbins),
                                  oracle.apply_threshold(f[0], method, nbins=nbins)), (method, nbins)
        assert np.array_equal(operations.apply_threshold(f[0], "otsu", nbins=64), oracle.apply_threshold(f[0], "otsu", nbins=64))
    u8 = (base[0] / 8).clip(0, 255).astype(np.uint8)
    assert np.array_equal(operations.apply_threshold(u8, method), oracle.apply_threshold(u8.copy(), method))
    assert not operations.apply_threshold(np.full((32, 32), 7, dtype=np.uint16), method).any()
    with pytest.raises(TypeError, match="unexpected keyword argument"):
        operations.apply_threshold(u16[0], method, window_size=15)


def test_li_float_building_blocks_against_numpy():
    """csrc/li.cu piece by piece: min(diff(unique(.))) through the device sort, and the stable split (= NumPy's
    boolean-mask indexing, order kept) whose pairwise sums equal np.sum of the compacted arrays bit for bit."""
    import ctypes as C

    import torch

    from arcadia_microscopy_tools_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(321)
    for n in (1, 2, 1000, 2048, 2049, 5000, 70001, 300000):
        x = rng.gamma(2.0, 0.1, n)
        x[rng.integers(0, n, n // 3)] = x[0]  # duplicates
        d = torch.from_numpy(x).cuda()
        nbytes = lib.amt_li_min_gap_scratch_bytes(n)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        gap = torch.empty(1, dtype=torch.float64, device="cuda")
        L.check(lib.amt_li_min_gap_f64(d.data_ptr(), n, gap.data_ptr(), scratch.data_ptr(), nbytes, None))
        uniq = np.unique(x)
        want = np.min(np.diff(uniq)) if uniq.size > 1 else np.inf
        assert gap.item() == want, (n, gap.item(), want)
        sorted_copy = scratch[: 8 * n].view(torch.float64).cpu().numpy()
        assert np.array_equal(sorted_copy, np.sort(x)), n
        t = float(np.median(x))
        above, rest = torch.empty_like(d), torch.empty_like(d)
        totals = torch.empty(2, dtype=torch.int64, device="cuda")
        nb2 = lib.amt_li_split_scratch_bytes(n)
        scratch2 = torch.empty(nb2, dtype=torch.uint8, device="cuda")
        L.check(lib.amt_li_split_f64(d.data_ptr(), n, C.c_double(t), above.data_ptr(), rest.data_ptr(), totals.data_ptr(),
                                     scratch2.data_ptr(), nb2, None))
        na, nr = totals.tolist()
        assert np.array_equal(above[:na].cpu().numpy(), x[x > t]) and np.array_equal(rest[:nr].cpu().numpy(), x[~(x > t)]), n
        if na:
            assert _gpu.plane_sums_f64(above[:na].reshape(1, na))[0] == np.sum(x[x > t])


@pytest.mark.parametrize("method", ["otsu", "li", "yen", "isodata", "mean", "minimum", "triangle"])
def test_thresholds_of_wide_integer_images(method):
    """int32 / int64 images (negative values included) are histogrammed value by value in scikit-image;
    here they travel as image - min in uint16 and the scans see the true bin centres."""
    rng = np.random.default_rng(78)
    shape = (120, 110)
    fg = rng.random(shape) < 0.3
    base = np.where(fg, rng.normal(900, 70, shape), rng.gamma(2.0, 70.0, shape))
    for dtype, shift in ((np.int32, -700), (np.int64, 100000), (np.int16, -300), (np.uint32, 0)):
        img = (base + shift).astype(dtype)
        got = operations.apply_threshold(img, method)
        want = oracle.apply_threshold(img.copy(), method)
        assert 0 < want.sum() < want.size and np.array_equal(got, want), (method, dtype)


@pytest.mark.parametrize("method", ["niblack", "sauvola"])
def test_window_thresholds_match_oracle(method):
    """box mean / standard deviation thresholds.  Integer window sums are exact, every float
    step is rounded once as in NumPy: threshold images and masks must equal the oracle's bit for bit."""
    import torch

    rng = np.random.default_rng(91)
    shape = (3, 101, 77)
    fg = rng.random(shape) < 0.3
    u16 = np.where(fg, rng.normal(9000, 700, shape), rng.gamma(2.0, 700.0, shape)).clip(0, 65535).astype(np.uint16)
    u16[1, :40, :30] = 65535  # saturated patch: variance clips at zero
    for kw in ({}, {"window_size": 3}, {"window_size": (5, 31), "k": 0.35}, {"window_size": 127, "k": -0.1}):
        for i in range(3):
            want = oracle.apply_threshold(u16[i], method, **kw)
            got = operations.apply_threshold(u16[i], method, **kw)
            assert got.dtype == np.bool_ and np.array_equal(got, want), (method, kw, i, int((got != want).sum()))
        got = operations.apply_threshold(u16, method, _batched=True, **kw)
        assert all(np.array_equal(got[i], oracle.apply_threshold(u16[i], method, **kw)) for i in range(3))
    # the threshold image itself (float64) is bit-identical
    func = oracle.threshold.threshold_niblack if method == "niblack" else oracle.threshold.threshold_sauvola
    d = torch.from_numpy(u16.view(np.int16)).cuda()
    _, thr = _gpu.window_threshold_u16(d, (15, 15), 1 if method == "sauvola" else 0, 0.2, 32767.5, want_thresholds=True)
    for i in range(3):
        _bits_equal(thr[i].cpu().numpy(), func(u16[i]), f"{method} threshold image {i}")
    if method == "sauvola":
        assert np.array_equal(operations.apply_threshold(u16[0], method, r=128.0), oracle.apply_threshold(u16[0], method, r=128.0))
    u8 = (u16[0] >> 8).astype(np.uint8)  # sauvola's default r follows the dtype: 127.5 for uint8
    assert np.array_equal(operations.apply_threshold(u8, method), oracle.apply_threshold(u8, method))
    assert not operations.apply_threshold(np.full((40, 40), 900, np.uint16), method).any()
    with pytest.raises(ValueError, match="must not be even"):
        operations.apply_threshold(u16[0], method, window_size=14)
    with pytest.raises(NotImplementedError, match="2\\*\\*53"):
        operations.apply_threshold(np.full((2048, 2048), 65535, np.uint16) - (np.arange(2048) % 2).astype(np.uint16), method)


@pytest.mark.parametrize("method", ["niblack", "sauvola"])
def test_window_thresholds_of_float_images_match_oracle(method):
    """a preprocessed plane is float64.  scikit-image's float route (np.pad 'reflect', float64 integral
    images by sequential np.cumsum along axis 0 then 1, four-corner sums in _correlate_sparse's order) is reproduced
    addition by addition: threshold images and masks equal the oracle's bit for bit."""
    import torch

    rng = np.random.default_rng(93)
    shape = (3, 101, 77)
    fg = rng.random(shape) < 0.3
    f64 = np.where(fg, rng.normal(0.4, 0.05, shape), rng.gamma(2.0, 0.02, shape))
    f64[1, :40, :30] = 1.0  # flat patch: g2 - m*m rounds to either side of zero and is clipped
    f64[2] -= 0.3  # negative values
    func = oracle.threshold.threshold_niblack if method == "niblack" else oracle.threshold.threshold_sauvola
    for kw in ({}, {"window_size": 3}, {"window_size": (5, 31), "k": 0.35}, {"window_size": 151, "k": -0.1}):
        win = kw.get("window_size", 15)
        if win == 151:
            planes = np.ascontiguousarray(np.tile(f64, (1, 2, 3)))  # a window above the integer path's 127 limit
        else:
            planes = f64
        for i in range(3):
            want = oracle.apply_threshold(planes[i].copy(), method, **kw)
            got = operations.apply_threshold(planes[i], method, **kw)
            assert got.dtype == np.bool_ and np.array_equal(got, want), (method, kw, i, int((got != want).sum()))
        got = operations.apply_threshold(planes, method, _batched=True, **kw)
        assert all(np.array_equal(got[i], oracle.apply_threshold(planes[i].copy(), method, **kw)) for i in range(3))
    d = torch.from_numpy(f64).cuda()
    _, thr = _gpu.window_threshold_f64(d, (15, 9), 1 if method == "sauvola" else 0, 0.2, 1.0, want_thresholds=True)
    for i in range(3):
        kw = {"r": 1.0} if method == "sauvola" else {}
        _bits_equal(thr[i].cpu().numpy(), func(f64[i].copy(), window_size=(15, 9), k=0.2, **kw), f"{method} float threshold image {i}")
    # the realistic chain: a rescaled DoG plane
    fov = make_fov(17, 2, 256, 320, n_cells=60)[0]
    pre = operations.rescale_by_percentile(operations.subtract_background_dog(fov[1]), (1, 99))
    assert np.array_equal(operations.apply_threshold(pre, method, window_size=25), oracle.apply_threshold(pre.copy(), method, window_size=25))
    assert not operations.apply_threshold(np.full((40, 40), 0.25), method).any()
    with pytest.raises(ValueError, match="must not be even"):
        operations.apply_threshold(f64[0], method, window_size=14)


def test_threshold_local_matches_oracle():
    """threshold_local (Gaussian-weighted mean, mode='reflect'): scipy's correlate1d order on the raw values."""
    rng = np.random.default_rng(92)
    shape = (2, 90, 123)
    fg = rng.random(shape) < 0.3
    base = np.where(fg, rng.normal(9000, 700, shape), rng.gamma(2.0, 700.0, shape)).clip(0, 65535)
    u16 = base.astype(np.uint16)
    f64 = base / 65535.0
    for kw in ({}, {"block_size": 35}, {"block_size": (5, 51), "offset": 120.0}, {"block_size": 21, "mode": "nearest"},
               {"block_size": 9, "param": 3.0, "offset": -40}):
        for i in range(2):
            assert np.array_equal(operations.apply_threshold(u16[i], "local", **kw), oracle.apply_threshold(u16[i], "local", **kw)), (kw, i)
        got = operations.apply_threshold(u16, "local", _batched=True, **kw)
        assert all(np.array_equal(got[i], oracle.apply_threshold(u16[i], "local", **kw)) for i in range(2))
    fkw = {"block_size": 35, "offset": 0.002}
    assert np.array_equal(operations.apply_threshold(f64[0], "local", **fkw), oracle.apply_threshold(f64[0], "local", **fkw))
    tiny = u16[0, :7, :5]  # the window reflects more than once around a tiny image
    assert np.array_equal(operations.apply_threshold(tiny, "local", block_size=35), oracle.apply_threshold(tiny, "local", block_size=35))
    with pytest.raises(ValueError, match="block_size must be odd"):
        operations.apply_threshold(u16[0], "local", block_size=10)
    with pytest.raises(NotImplementedError, match="mode 'wrap'"):
        operations.apply_threshold(u16[0], "local", block_size=11, mode="wrap")


def test_contracted_dog_mode_stays_within_tolerance_and_is_opt_in():
    """amt_tune('dog_fma', 1): fused multiply-adds in the DoG (2 DP instructions per tap pair instead of 3).
    Opt-in; the planes then differ from scipy's only in the last bits (tolerance 1e-12 relative to the plane's
    scale here, 1e-5 in the north star), thresholds and masks of the test images are unchanged, and switching
    it off restores bit-exactness."""
    from arcadia_microscopy_tools_b200 import _lib

    lib = _lib.load()
    fov = make_fov(11, 2, 512, 768, n_cells=150)[0]
    want = oracle.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0)
    exact = operations.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0)
    _bits_equal(exact, want, "exact mode")
    _lib.check(lib.amt_tune(b"dog_fma", 1))
    try:
        fast = operations.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0)
        pipe = Pipeline([ImageOperation(operations.subtract_background_dog, 0.6, 16.0, percentile=0),
                         ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99)),
                         ImageOperation(operations.apply_threshold)])
        fast_mask = pipe(fov[1])
    finally:
        _lib.check(lib.amt_tune(b"dog_fma", 0))
    assert np.max(np.abs(fast - want)) <= 1e-12 * np.max(np.abs(want))
    assert not np.array_equal(fast, want)  # it really is a different rounding
    exact_mask = oracle.apply_threshold(oracle.rescale_by_percentile(want, (1, 99)))
    assert np.array_equal(fast_mask, exact_mask)
    _bits_equal(operations.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0), want, "exact mode restored")


def test_bucketed_selection_matches_sorted_order():
    """amt_select_f64_bucketed (2-byte monotone buckets + sparse value gather, the executor's percentile
    path) must return exactly the order statistics np.sort gives, also for distributions that put most
    samples in one bucket (whole-plane radix fallback), signed zeros and values outside the bucket range."""
    import ctypes as C

    import torch

    from arcadia_microscopy_tools_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(123)
    n = 512 * 512
    planes = {
        "dog_like": rng.normal(0, 1.5e-3, n) + (rng.random(n) < 0.08) * rng.gamma(2.0, 0.01, n),
        "narrow": 0.25 + rng.normal(0, 1e-9, n),                      # one bucket holds everything
        "zeros": np.where(rng.random(n) < 0.6, 0.0, rng.normal(0, 1e-4, n)) * rng.choice([-1.0, 1.0], n),
        "wide": rng.normal(0, 1.0, n) * 10.0 ** rng.integers(-30, 6, n),  # far beyond 2^-20 .. 2^12
        "tiny": rng.normal(0, 1e-9, n),                                # all below the finest bucket
    }
    data = np.stack(list(planes.values()))
    ranks = [0, 1, n // 100, n // 100 + 1, n // 2, (99 * n) // 100, n - 2, n - 1]
    d = torch.from_numpy(data).cuda()
    buckets = torch.empty(data.shape, dtype=torch.int16, device="cuda")
    L.check(lib.amt_bucket12(_gpu.ptr(d), _gpu.ptr(buckets), data.size, _gpu.stream_ptr()))
    b = buckets.cpu().numpy().view(np.uint16)
    assert b.max() < 4096
    for i in range(data.shape[0]):  # buckets are monotone in the value (-0.0 sits one bucket below +0.0)
        order = np.argsort(data[i], kind="stable")
        v, bb = data[i][order], b[i][order].astype(np.int64)
        both_zero = (v[1:] == 0) & (v[:-1] == 0)
        assert np.all((np.diff(bb) >= 0) | both_zero), list(planes)[i]
    mm = _gpu.minmax_keys(d)
    out = torch.empty((data.shape[0], len(ranks)), dtype=torch.float64, device="cuda")
    nbytes = lib.amt_select_f64_scratch_bytes(data.shape[0], n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    r = (C.c_int64 * len(ranks))(*ranks)
    L.check(lib.amt_select_f64_bucketed(_gpu.ptr(d), _gpu.ptr(buckets), data.shape[0], n, r, len(ranks), _gpu.ptr(mm),
                                        _gpu.ptr(out), _gpu.ptr(scratch), nbytes, _gpu.stream_ptr()))
    got = out.cpu().numpy()
    want = np.sort(data, axis=1)[:, ranks]
    assert np.array_equal(got, want), [(list(planes)[i], got[i], want[i]) for i in range(len(planes)) if not np.array_equal(got[i], want[i])][:2]


def test_device_chain_after_crop_and_narrow_dtypes():
    """A device-resident Pipeline hands each operation the previous one's tensor: a centred crop is a strided view
    (it must be packed before a kernel sees it), and a uint8 / int32 / bool first input keeps its own dtype rules
    (img_as_float's 1/255 for uint8, (2x + 1) / (max - min) for signed integers)."""
    rng = np.random.default_rng(321)
    img = rng.integers(100, 5000, size=(3, 150, 170)).astype(np.uint16)
    crop = ImageOperation(operations.crop_to_center, (96, 112))
    dog = ImageOperation(operations.subtract_background_dog, low_sigma=0.6, high_sigma=4.0, percentile=5)
    rescale = ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99))

    def want_for(plane):
        c = oracle.crop_to_center(plane, (96, 112))
        return oracle.rescale_by_percentile(oracle.subtract_background_dog(c, 0.6, 4.0, 5), (1, 99))

    got = Pipeline([crop, dog, rescale])(img[0])
    _bits_equal(got, want_for(img[0]), "crop -> dog -> rescale")
    got = Pipeline([crop, dog, rescale], parallel=True)(img)
    for i in range(3):
        _bits_equal(got[i], want_for(img[i]), f"parallel crop chain, slice {i}")
    dev = operations.crop_to_center(_gpu.to_device(img[1]), (96, 112))
    assert not dev.is_contiguous()
    _bits_equal(_gpu.to_host(operations.subtract_background_dog(dev, 0.6, 4.0, 5)),
                oracle.subtract_background_dog(oracle.crop_to_center(img[1], (96, 112)), 0.6, 4.0, 5), "sliced device input")
    u8 = (img[0] >> 5).astype(np.uint8)
    _bits_equal(Pipeline([dog, rescale])(u8), oracle.rescale_by_percentile(oracle.subtract_background_dog(u8, 0.6, 4.0, 5), (1, 99)),
                "uint8 device chain")
    i32 = img[0].astype(np.int32) - 2000
    _bits_equal(operations.subtract_background_dog(i32, 0.6, 4.0, 5), oracle.subtract_background_dog(i32, 0.6, 4.0, 5), "int32 DoG")
    _bits_equal(operations.gaussian_smooth(img[0].astype(np.uint32), 2.0), filters.gaussian(img[0].astype(np.uint32), 2.0), "uint32 gaussian")
