"""Host half of the run-length staging of label masks (``amt_rle_encode_host``, csrc/host_rle.cpp): no GPU involved.

The encoder turns every row of a label mask into runs {value, end column}; decoding them with a plain Python loop must
give the mask back (with negative labels as background and values beyond int32 saturated, as on the device routes)."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from arcadia_microscopy_tools_b200 import _lib
from arcadia_microscopy_tools_b200.synthetic import make_fov

DTYPES = {np.dtype(np.int64): _lib.AMT_I64, np.dtype(np.int32): _lib.AMT_I32, np.dtype(np.uint16): _lib.AMT_U16}


def encode(lab: np.ndarray, threads: int):
    lib = _lib.load()
    n, H, W = lab.shape
    runs = np.zeros((n * H * (W // 4), 2), np.uint32)
    rows = np.zeros((n * H, 2), np.uint32)
    neg = np.zeros(n, np.int32)
    n_runs = C.c_int64(0)
    rc = lib.amt_rle_encode_host(lab.ctypes.data, DTYPES[lab.dtype], n, H, W, threads, runs.ctypes.data, rows.ctypes.data,
                                 neg.ctypes.data, C.byref(n_runs))
    return rc, runs, rows, neg, n_runs.value


def decode(runs, rows, n, H, W):
    out = np.zeros((n * H, W), np.int64)
    for r in range(n * H):
        first, count = (int(v) for v in rows[r])
        x = 0
        for k in range(first, first + count):
            v, end = (int(t) for t in runs[k])
            assert end > x
            out[r, x:end] = v
            x = end
        assert x == W
    return out.reshape(n, H, W)


@pytest.fixture(scope="module")
def masks():
    g = make_fov(1, 1, 1024, 1024, 300)[1]
    return np.stack([g, g[::-1].copy()]).astype(np.int64)  # 2 x 1024 x 1024: large enough for the threads to split it


@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.uint16])
@pytest.mark.parametrize("threads", [1, 3])
def test_runs_decode_to_the_mask(masks, dtype, threads):
    lab = masks.astype(dtype)
    rc, runs, rows, neg, n_runs = encode(lab, threads)
    assert rc == _lib.AMT_OK and not neg.any()
    assert int(rows[:, 1].sum()) == n_runs and n_runs < lab.size // 20
    assert np.array_equal(decode(runs, rows, *lab.shape), masks)


def test_negative_and_huge_labels(masks):
    lab = masks.copy()
    lab[1, 5, 7] = -3
    lab[0, 0, 0] = 2**40
    lab[0, -1, -1] = 70000
    rc, runs, rows, neg, _ = encode(lab, 2)
    want = lab.copy()
    want[1, 5, 7] = 0
    want[0, 0, 0] = 2**31 - 1
    assert rc == _lib.AMT_OK and list(neg) == [0, 1]
    assert np.array_equal(decode(runs, rows, *lab.shape), want)


def test_widths_that_are_no_multiple_of_the_vector_step():
    rng = np.random.default_rng(5)
    for W in (16, 17, 31, 100, 257):
        lab = np.repeat(rng.integers(0, 9, (1, 40, W // 5 + 1)), 5, axis=2)[:, :, :W].astype(np.uint16)
        lab = np.ascontiguousarray(lab)
        rc, runs, rows, _, _ = encode(lab, 1)
        assert rc == _lib.AMT_OK, W
        assert np.array_equal(decode(runs, rows, *lab.shape), lab), W


def test_ragged_mask_does_not_fit():
    noise = np.random.default_rng(0).integers(0, 50, (1, 256, 256)).astype(np.int32)
    assert encode(noise, 2)[0] == _lib.AMT_ERR_CAPACITY
