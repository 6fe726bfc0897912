"""Throughput of the other BASELINE.json configurations on one GPU (bench.py measures config 2):

  config3  time-lapse, T x 2 channels x 2048 x 2048: every frame preprocessed (both channels) and the
           segmentation channel thresholded + labelled + quantified, through the batch executor
  config4  confocal z-stack, 64 x 4 x 1024 x 1024: per-slice preprocessing + 2-D threshold labelling of
           every slice through the batch executor, then 3-D per-object quantification of a label volume
           over the four raw channels (volumes.quantify_label_volume)

Inputs are synthetic and resident in HBM; times are CUDA-event (executor) / synchronised wall clock
(3-D table).  Prints one JSON line per configuration.  Measurement tooling, not part of the library.

    python scripts/bench_configs.py [--frames 500] [--steps 3]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from arcadia_microscopy_tools_b200 import _gpu  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402
from arcadia_microscopy_tools_b200.synthetic import BACKGROUND_LAMBDA, CHANNEL_GAIN, make_cell_layer  # noqa: E402
from arcadia_microscopy_tools_b200.volumes import quantify_label_volume  # noqa: E402


def device_frames(n, c, h, w, n_cells, dev, seed):
    layers = [torch.from_numpy(make_cell_layer(seed + u, h, w, n_cells)[0]).to(dev) for u in range(4)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    out = torch.empty((n, c, h, w), dtype=torch.int16, device=dev)
    lam = torch.empty((h, w), dtype=torch.float32, device=dev)
    for i in range(n):
        for ch in range(c):
            lam.fill_(BACKGROUND_LAMBDA[ch % 4])
            img = torch.poisson(lam, generator=gen).to(torch.float64) + CHANNEL_GAIN[ch % 4] * layers[i % 4]
            v = img.round_().clamp_(0, 65535).to(torch.int32)
            out[i, ch] = torch.where(v >= 32768, v - 65536, v).to(torch.int16)
    torch.cuda.synchronize()
    return out


def run_executor(frames, seg_channel, steps, chunk):
    n, c, h, w = frames.shape
    cfg = FovPipelineConfig(n_channels=c, height=h, width=w, seg_channel=seg_channel, chunk_fovs=chunk, max_labels=4096,
                            max_label_value=0, quantify_given_mask=False)
    with FovBatchExecutor(cfg, device=0) as ex:
        out = ex.alloc_outputs(n)
        for _ in range(3):
            ex.run_device(frames, None, out)
        ms = [ex.run_device(frames, None, out) for _ in range(steps)]
        cells = float(out["counts_thr"].float().mean().item())
    return float(np.median(ms)), cells


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)

    # ---- config 3
    T, C, H, W = args.frames, 2, 2048, 2048
    frames = device_frames(T, C, H, W, 2000, dev, 31000)
    ms, cells = run_executor(frames, 0, args.steps, 16)
    bytes_px = C * 34 + 20 + (4 + 2 * C)  # SURVEY 8(d) accounting: stage A per channel, B once, C for one mask
    print(json.dumps({"config": f"config3: T={T} x {C} x {H}x{W} uint16, per-frame W_pre (both channels) + Otsu/CCL + tables",
                      "ms_per_pass": ms, "frames_per_s": T / ms * 1e3, "mpix_per_s": T * C * H * W / ms / 1e3,
                      "algorithmic_gbs": T * H * W * bytes_px / ms / 1e6, "cells_per_frame": cells}), flush=True)
    del frames
    torch.cuda.empty_cache()

    # ---- config 4
    Z, C, H, W = 64, 4, 1024, 1024
    stack = device_frames(Z, C, H, W, 500, dev, 41000)
    ms, cells = run_executor(stack, 1, args.steps, 16)
    rng = np.random.default_rng(41)
    vol = np.zeros((Z, H, W), dtype=np.int32)
    zz = np.arange(Z)[:, None, None]
    k = 0
    for _ in range(2000):  # ~2k ellipsoids, drawn into their bounding boxes
        cz, cy, cx = rng.uniform([0, 0, 0], [Z, H, W])
        rz, ry, rx = rng.uniform([2, 6, 6], [6, 14, 14])
        z0, z1 = max(int(cz - rz), 0), min(int(cz + rz) + 2, Z)
        y0, y1 = max(int(cy - ry), 0), min(int(cy + ry) + 2, H)
        x0, x1 = max(int(cx - rx), 0), min(int(cx + rx) + 2, W)
        g = np.mgrid[z0:z1, y0:y1, x0:x1]
        m = ((g[0] - cz) / rz) ** 2 + ((g[1] - cy) / ry) ** 2 + ((g[2] - cx) / rx) ** 2 <= 1
        sub = vol[z0:z1, y0:y1, x0:x1]
        if m.any() and not (sub[m] > 0).any():
            k += 1
            sub[m] = k
    del zz
    d_vol = _gpu.to_device(vol)
    chans = {n: stack[:, i].contiguous() for i, n in enumerate(["brightfield", "dapi", "fitc", "tritc"])}
    for _ in range(2):
        table = quantify_label_volume(d_vol, chans)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        table = quantify_label_volume(d_vol, chans)
    torch.cuda.synchronize()
    ms3d = (time.perf_counter() - t0) / args.steps * 1e3
    print(json.dumps({"config": f"config4: Z={Z} x {C} x {H}x{W} uint16, per-slice W_pre + Otsu/CCL + tables, then 3-D tables of {k} objects",
                      "ms_per_slice_pass": ms, "slices_per_s": Z / ms * 1e3, "mpix_per_s": Z * C * H * W / ms / 1e3,
                      "cells_per_slice": cells, "ms_3d_quantification": ms3d, "objects": int(len(table["label"])),
                      "voxel_gbs_3d": Z * H * W * (4 + 2 * C) / ms3d / 1e6}), flush=True)


if __name__ == "__main__":
    main()
