#!/usr/bin/env python
"""Benchmark of the preprocess -> threshold/label -> quantify hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of workload W over one batch of synthetic FOVs per GPU
(config 2 of BASELINE.json: 256 FOVs x 4 channels x 2048 x 2048 uint16 + one ~2k-cell integer
label mask per FOV).  FOVs are independent, so every rank runs its own batch with no
collective on the data path (weak scaling); torch.distributed is used only for the barrier
and the max-over-ranks time.  Prints ONE JSON line on rank 0.

`value`   : Mpix/s (input samples: FOVs x C x H x W per second), inputs resident in HBM,
            timed with CUDA events on the executor's compute stream.
`e2e`     : same metric through the host-fed C-ABI call (amt_executor_run_host) with pinned
            host buffers; H2D of every FOV and label mask and D2H of the per-cell tables are
            inside the timed region.
`roofline`: the dominant kernel (the sigma=16 float64 Gaussian pass) timed alone with CUDA
            events; achieved = its compulsory bytes / time vs the measured HBM peak, plus the
            FP64-pipe view that actually bounds it (`roofline_fp64`).
`cpu_baseline`: the oracle (NumPy/SciPy restatement of the reference's call chain) timed on the
            host cores for a bounded sample (N=1 only).
`--impl reference`: the same CPU implementation on all host cores (one FOV per worker process).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C, H, W = 4, 2048, 2048
N_CELLS = 2000
SEG_CHANNEL = 1
ALGO_BYTES_PER_FOV_PIXEL = 180  # 4*34 + 20 + 2*(4 + 2*4)
METRIC = "Mpix/s preprocess+label+quantify (input samples per second)"


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()) | {"source": "measured"}
    return {"hbm_gbs": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs: NVML in-process every 10 ms
    (nvidia-ml-py), or `nvidia-smi` every 200 ms when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int) -> None:
        self.index = index
        self.sm: list[float] = []
        self.sm_max: float | None = None
        self.reasons: set[str] = set()
        self.source = "nvidia-smi"
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in visible.split(",") if v.strip().isdigit()]
            physical = int(ids[index]) if index < len(ids) else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self) -> None:
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
        try:
            bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
        except Exception:
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
        for name, bit in self.REASON_BITS.items():
            if bits & bit:
                self.reasons.add(name)

    def _sample_smi(self) -> None:
        out = subprocess.run(
            ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
            capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.splitlines()[0].split(",")]
        if r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if len(r) > 1 and r[1].replace(".", "").isdigit():
            self.sm_max = max(self.sm_max or 0.0, float(r[1]))
        for k, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                self.reasons.add(name)

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        order = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": [n for n in order if n in self.reasons], "samples": len(self.sm), "source": self.source}


# ------------------------------------------------------------------------------------ inputs
def build_device_batch(n_fov: int, n_unique: int, device, seed0: int = 20260000):
    """(n_fov, C, H, W) uint16 bits + (n_fov, H, W) int32 labels, resident in HBM.

    `n_unique` seeded cell layouts (NumPy, SURVEY 8d generator) are rendered on the host; every
    one of the n_fov slots gets its own Poisson background drawn on the GPU (seeded torch
    generator), so all slots hold distinct data and separate memory."""
    import torch

    from arcadia_microscopy_tools_b200.synthetic import BACKGROUND_LAMBDA, CHANNEL_GAIN, make_cell_layer

    layers, labels = [], []
    for u in range(n_unique):
        layer, lab, _ = make_cell_layer(seed0 + u, H, W, N_CELLS)
        layers.append(torch.from_numpy(layer).to(device))
        labels.append(torch.from_numpy(lab).to(device))
    gen = torch.Generator(device=device)
    gen.manual_seed(seed0)
    fovs = torch.empty((n_fov, C, H, W), dtype=torch.int16, device=device)
    given = torch.empty((n_fov, H, W), dtype=torch.int32, device=device)
    lam = torch.empty((H, W), dtype=torch.float32, device=device)
    for i in range(n_fov):
        layer = layers[i % n_unique]
        given[i] = labels[i % n_unique]
        for c in range(C):
            lam.fill_(BACKGROUND_LAMBDA[c])
            img = torch.poisson(lam, generator=gen).to(torch.float64) + CHANNEL_GAIN[c] * layer
            v = img.round_().clamp_(0, 65535).to(torch.int32)
            fovs[i, c] = torch.where(v >= 32768, v - 65536, v).to(torch.int16)  # uint16 bit pattern
    torch.cuda.synchronize(device)
    return fovs, given, int(max(int(l.max()) for l in labels))


# ------------------------------------------------------------------------------------ CPU side
def oracle_fov(args) -> int:
    """Workload W for one FOV through the oracle (the reference's call chain on SciPy/NumPy)."""
    fov, given = args
    import oracle

    names = ["brightfield", "dapi", "fitc", "tritc"]
    morph = ["label", "area", "bbox", "centroid", "axis_major_length", "axis_minor_length", "eccentricity", "orientation"]
    inten = ["intensity_mean", "intensity_max", "intensity_min", "intensity_std"]
    pre = [oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[c], 0.6, 16.0, 0), (1, 99), (0, 1))
           for c in range(fov.shape[0])]
    mask = oracle.apply_threshold(pre[SEG_CHANNEL])
    chans = {n: fov[i] for i, n in enumerate(names[: fov.shape[0]])}
    total = 0
    for m in (mask, given.astype(np.int64)):
        labels = oracle.process_mask(m, True)
        props = oracle.cell_properties(labels, chans, morph, inten)
        total += len(props["label"])
    return total


def host_fov(seed: int):
    from arcadia_microscopy_tools_b200.synthetic import make_fov

    fov, given, _ = make_fov(seed, C, H, W, N_CELLS)
    return fov, given


def cpu_baseline_single(n_fovs: int = 3) -> dict:
    batch = [host_fov(20260000 + i) for i in range(n_fovs)]
    t0 = time.perf_counter()
    for item in batch:
        oracle_fov(item)
    dt = time.perf_counter() - t0
    return {"value": n_fovs * C * H * W / dt / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"{n_fovs} FOVs of {C}x{H}x{W} uint16 + their label masks, full workload W, single thread (oracle: "
                      "reference call chain on scipy/numpy; scikit-image itself is not installable here)"}


def ncu_traffic(dom: str, planes: int) -> tuple[float | None, str | None]:
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r01_ncu_full_dog_strip.json, same 32-plane launch as the live timing)."""
    path = ROOT / "profiles" / "r01_ncu_full_dog_strip.json"
    if not path.exists() or planes != 32:
        return None, None
    want = "unsigned short" if dom == "axis0" else "double, 1"
    for k in json.loads(path.read_text())["kernels"]:
        if want in k["kernel"] and "dram_bytes" in k:
            return float(k["dram_bytes"]), f"profiles/{path.name} ({k['kernel']})"
    return None, None


def run_reference(args) -> None:
    """--impl reference: the CPU implementation on all host cores, one FOV per worker process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))  # every host core (each worker holds ~1 GB of float64 planes)
    batch = [host_fov(20260000 + i % 2) for i in range(workers)]
    times = []
    with mp.get_context("fork").Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(oracle_fov, batch)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = len(times) * workers * C * H * W / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "fov_per_s": len(times) * workers / total,
        "config": {"workload": f"config2: {C}x{H}x{W} uint16 FOVs + ~{N_CELLS}-cell label mask, workload W; "
                               f"bounded sample of {workers} FOVs per step"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": workers, "kind": "port",
                         "sample": f"{workers} FOVs per step, one worker process per FOV (oracle port of the reference chain)"},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------ GPU side
def time_kernels(lib, gpu, fovs, cfg_hw, steps: int, warmup: int, exact_every: int = 0, exact_offset: int = 0) -> dict:
    """Per-kernel CUDA-event timings of the two Gaussian passes and the FP64 probe.  exact_every > 0: the launch
    the executor makes by default (plane p keeps scipy's operation order iff p % exact_every == exact_offset, the
    other planes use fused multiply-adds); 0: every plane in scipy's order."""
    import ctypes as CT

    import torch

    from arcadia_microscopy_tools_b200 import _lib as L

    hw_lo, hw_hi = cfg_hw
    n_fov = min(8, fovs.shape[0])
    planes = n_fov * C
    dev = fovs.device
    d_lo = torch.from_numpy(hw_lo).to(dev)
    d_hi = torch.from_numpy(hw_hi).to(dev)
    tmp_lo = torch.empty((planes, H, W), dtype=torch.float64, device=dev)
    tmp_hi = torch.empty_like(tmp_lo)
    out = torch.empty_like(tmp_lo)
    mm = torch.empty((planes, 2), dtype=torch.int64, device=dev)
    probe_scratch = torch.empty(148 * 8 * 256, dtype=torch.float64, device=dev)
    st = gpu.stream_ptr()

    def v_pass():
        L.check(lib.amt_dog2d_axis0(gpu.ptr(fovs), L.AMT_U16, 1.0 / 65535.0, planes, H, W, gpu.ptr(d_lo), len(hw_lo) - 1,
                                    gpu.ptr(d_hi), len(hw_hi) - 1, gpu.ptr(tmp_lo), gpu.ptr(tmp_hi), st))

    def h_pass():
        L.check(lib.amt_dog2d_axis1(gpu.ptr(tmp_lo), gpu.ptr(tmp_hi), gpu.ptr(out), planes, H, W, gpu.ptr(d_lo),
                                    len(hw_lo) - 1, gpu.ptr(d_hi), len(hw_hi) - 1, gpu.ptr(mm), st))

    dp = CT.c_uint64(0)

    def probe():
        L.check(lib.amt_fp64_probe(4096, gpu.ptr(probe_scratch), CT.byref(dp), st))

    def timed(fn) -> float:
        for _ in range(max(warmup, 3)):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    L.check(lib.amt_tune(b"dog_exact_every", exact_every), "amt_tune")
    L.check(lib.amt_tune(b"dog_exact_offset", exact_offset), "amt_tune")
    try:
        ms_v, ms_h = timed(v_pass), timed(h_pass)
    finally:
        L.check(lib.amt_tune(b"dog_exact_every", 0), "amt_tune")
        L.check(lib.amt_tune(b"dog_exact_offset", 0), "amt_tune")
    ms_p = timed(probe)
    px = planes * H * W
    r_lo, r_hi = len(hw_lo) - 1, len(hw_hi) - 1
    exact_share = 1.0 / exact_every if exact_every > 0 else 1.0
    # DP instructions per sample and axis: DADD + DMUL + DADD per tap pair in scipy's order, DADD + DFMA contracted
    dp_per_px_axis = exact_share * ((1 + 3 * r_lo) + (1 + 3 * r_hi)) + (1.0 - exact_share) * ((1 + 2 * r_lo) + (1 + 2 * r_hi))
    return {
        "planes": planes, "ms_axis0": ms_v, "ms_axis1": ms_h, "ms_probe": ms_p,
        "fp64_peak_tinstr_s": dp.value / (ms_p * 1e-3) / 1e12,
        "axis0": {"bytes": px * (2 + 16), "dp_instr": px * (dp_per_px_axis + 1)},
        "axis1": {"bytes": px * (16 + 8), "dp_instr": px * (dp_per_px_axis + 1)},
    }


def bind_near_gpu(local: int) -> list[int]:
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned staging it
    allocates (first touch) and the copy-issuing thread live on the GPU's own NUMA node: on an
    8-GPU box the host-fed path is bounded by host memory / PCIe root-complex locality."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        near = {i for i in range(n_cpu) if (mask[i // 64] >> (i % 64)) & 1}
        allowed = near & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)
    except Exception:
        pass
    return []


def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    from arcadia_microscopy_tools_b200 import _gpu, _lib
    from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cpus = bind_near_gpu(local) if world > 1 else []
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    n_fov = args.fovs
    fovs, given, max_label = build_device_batch(n_fov, args.unique, dev, seed0=20260000 + 1000 * rank)
    cfg = FovPipelineConfig(n_channels=C, height=H, width=W, seg_channel=SEG_CHANNEL, chunk_fovs=args.chunk,
                            max_labels=4096, max_label_value=max_label, given_label_dtype=np.uint16,
                            exact_all_channels=args.exact_all_channels)
    ex = FovBatchExecutor(cfg, device=local)
    out = ex.alloc_outputs(n_fov)

    # ---- device-resident timed region: W warm-up steps, then exactly K timed steps
    for _ in range(args.warmup):
        ex.run_device(fovs, given, out, sync=True)
    barrier()
    launches0 = lib.amt_launch_count()
    step_ms = []
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_ms.append(ex.run_device(fovs, given, out, sync=True))
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
    barrier()
    launches = int(lib.amt_launch_count() - launches0)
    dev_s = sum(step_ms) / 1e3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t[0]), float(t[1])
    counts = _gpu.to_host(out["counts_thr"]), _gpu.to_host(out["counts_given"])

    # ---- end to end through the host-fed C-ABI call (pinned host buffers)
    e2e = None
    if not args.no_e2e:
        # pinned host staging is per rank: halve it on multi-GPU runs so that 8 ranks stay well inside host RAM
        n_e2e = min(n_fov, args.e2e_fovs if world == 1 else min(args.e2e_fovs, 128))
        # label masks travel as uint16 (Cellpose's own mask dtype below 65536 cells): 8.4 instead of 16.8 MB per FOV
        if args.host_alloc == "torch":
            h_fovs = torch.empty((n_e2e, C, H, W), dtype=torch.int16, pin_memory=True)
            h_given = torch.empty((n_e2e, H, W), dtype=torch.int16, pin_memory=True)
            h_fovs.copy_(fovs[:n_e2e])
            h_given.copy_(given[:n_e2e].to(torch.int16))
            np_fovs = h_fovs.numpy().view(np.uint16)
            np_given = h_given.numpy().view(np.uint16)
        else:  # the library's own pinned staging (amt_host_alloc), optionally write-combined
            pin_fovs = _gpu.PinnedBuffer((n_e2e, C, H, W), np.uint16, write_combined=args.host_alloc == "wc")
            pin_given = _gpu.PinnedBuffer((n_e2e, H, W), np.uint16, write_combined=args.host_alloc == "wc")
            np_fovs, np_given = pin_fovs.array, pin_given.array
            np_fovs[...] = fovs[:n_e2e].cpu().numpy().view(np.uint16)
            np_given[...] = given[:n_e2e].to(torch.int16).cpu().numpy().view(np.uint16)
        h_out = ex.alloc_host_outputs(n_e2e)
        for _ in range(min(args.warmup, 2)):
            ex.run_host(np_fovs, np_given, h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ex.run_host(np_fovs, np_given, h_out)
        e2e_s = time.perf_counter() - t0
        barrier()
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        assert np.array_equal(h_out["counts_thr"], counts[0][:n_e2e]) and np.array_equal(h_out["counts_given"], counts[1][:n_e2e])
        d2h = sum(int(h_out[k].nbytes) for k in h_out)
        e2e = {"value": world * args.steps * n_e2e * C * H * W / e2e_s / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(np_fovs.nbytes + np_given.nbytes), "d2h_bytes_per_step": d2h,
               "fov_per_s": world * args.steps * n_e2e / e2e_s, "fovs_per_step": n_e2e, "timing": "wall clock around the synchronous C-ABI call",
               "rank0_cpu_affinity": f"{len(cpus)} CPUs local to the GPU (NVML)" if cpus else "unchanged",
               "host_buffers": {"torch": "torch pinned tensors", "pinned": "amt_host_alloc", "wc": "amt_host_alloc, write-combined"}[args.host_alloc]}

    # ---- the same device-resident pass with the DoG's multiply-adds contracted (amt_tune "dog_fma"): what scipy's
    # exact operation order costs.  Not the reported value: the default path stays bit-identical to the reference.
    contracted = strict = None
    if not args.no_contracted:
        import dataclasses

        keep = {k: out[k].clone() for k in ("tables_thr", "counts_thr", "thresholds")}

        def same_as_default() -> bool:
            return all(torch.equal(out[k].view(torch.int64) if out[k].dtype == torch.float64 else out[k],
                                   keep[k].view(torch.int64) if keep[k].dtype == torch.float64 else keep[k]) for k in keep)

        def timed_pass(executor) -> float:
            executor.run_device(fovs, given, out, sync=True)
            barrier()
            ms = [executor.run_device(fovs, given, out, sync=True) for _ in range(args.steps)]
            barrier()
            t_pass = torch.tensor([sum(ms) / 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_pass, op=dist.ReduceOp.MAX)
            return float(t_pass[0])

        _lib.check(lib.amt_tune(b"dog_fma", 1), "amt_tune")
        try:
            c_s = timed_pass(ex)
        finally:
            _lib.check(lib.amt_tune(b"dog_fma", 0), "amt_tune")
        contracted = {"value": world * args.steps * n_fov * C * H * W / c_s / 1e6, "unit": "Mpix/s",
                      "ms_per_step": 1e3 * c_s / args.steps,
                      "thresholds_counts_and_tables_bit_identical_to_default": same_as_default(),
                      "note": "opt-in amt_tune('dog_fma', 1): fused multiply-adds for EVERY channel, the segmentation channel "
                              "included (its plane then differs from scipy's in the last bits); NOT the reported value"}
        if not cfg.exact_all_channels:
            # the strict configuration: scipy's exact operation order for every channel (all float planes bit-identical)
            with FovBatchExecutor(dataclasses.replace(cfg, exact_all_channels=True), device=local) as ex_strict:
                s_s = timed_pass(ex_strict)
            strict = {"value": world * args.steps * n_fov * C * H * W / s_s / 1e6, "unit": "Mpix/s",
                      "ms_per_step": 1e3 * s_s / args.steps,
                      "thresholds_counts_and_tables_bit_identical_to_default": same_as_default(),
                      "note": "FovPipelineConfig(exact_all_channels=True): the float planes of the channels that are not "
                              "thresholded are bit-identical to scipy's too (default: equal to ~1e-15 relative)"}
        del keep

    # ---- per-kernel roofline (rank 0) and CPU baseline (rank 0, N=1)
    line = None
    if rank == 0:
        peaks = measured_peaks()
        hw = (_gpu.gaussian_half_weights(cfg.low_sigma), _gpu.gaussian_half_weights(cfg.high_sigma))
        mixed = not cfg.exact_all_channels and C > 1
        k = time_kernels(lib, _gpu, fovs, hw, steps=max(args.steps, 5), warmup=args.warmup,
                         exact_every=C if mixed else 0, exact_offset=SEG_CHANNEL if mixed else 0)
        k_exact = time_kernels(lib, _gpu, fovs, hw, steps=max(args.steps, 5), warmup=args.warmup) if mixed else k
        dom = "axis1" if k["ms_axis1"] >= k["ms_axis0"] else "axis0"
        dom_ms = k["ms_" + dom]
        achieved = k[dom]["bytes"] / (dom_ms * 1e-3) / 1e9
        fp64_ach = k[dom]["dp_instr"] / (dom_ms * 1e-3) / 1e12
        traffic, traffic_src = ncu_traffic(dom, k["planes"])
        dp_per_sample_axis = k[dom]["dp_instr"] / (k["planes"] * H * W) - 1
        ms_per_step = 1e3 * dev_s / args.steps
        value = world * args.steps * n_fov * C * H * W / dev_s / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"config2: {n_fov} FOVs/GPU x {C}x{H}x{W} uint16 + ~{N_CELLS}-cell int32 label mask per FOV; "
                                   "W = DoG(0.6,16)+pct rescale on 4 channels, Otsu+CCL+clear_border on ch1, per-cell tables "
                                   "for the threshold mask and the given mask",
                       "arithmetic": ("float64, scipy's exact operation order for every channel" if cfg.exact_all_channels else
                                      "float64; scipy's exact operation order for the segmentation channel (labels, counts, tables "
                                      "bit-exact by construction), fused multiply-adds for the other channels' float planes "
                                      "(equal to scipy's to ~1e-15 relative; tolerance 1e-5)"),
                       "fovs_per_gpu": n_fov, "unique_cell_layouts": args.unique, "chunk_fovs": args.chunk,
                       "l2_policy": f"inputs {fovs.numel() * 2 / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
                       "sharding": "FOV-independent, one process per GPU, no collective on the data path"},
            "fov_per_s": world * args.steps * n_fov / dev_s,
            "hbm_algorithmic": {"bytes_per_fov": ALGO_BYTES_PER_FOV_PIXEL * H * W,
                                "achieved_gbs": world * args.steps * n_fov * ALGO_BYTES_PER_FOV_PIXEL * H * W / dev_s / 1e9,
                                "frac_of_peak_per_gpu": args.steps * n_fov * ALGO_BYTES_PER_FOV_PIXEL * H * W / dev_s / 1e9 / peaks["hbm_gbs"]},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "gpu_launches": launches,
            "cells_per_fov": {"threshold_mask": float(counts[0].mean()), "given_mask": float(counts[1].mean())},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "exact_all_channels_mode": strict,
            "contracted_mode": contracted,
            "roofline": {"kernel": f"dog_strip_kernel ({dom} pass of the DoG: sigma 0.6 and 16 filters of {k['planes']} planes)",
                         "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": k[dom]["bytes"], "peak_source": peaks["source"],
                         "ms_per_launch": dom_ms,
                         "note": "this kernel is FP64-pipe-bound by construction (193 DP instr per sample and axis at sigma=16 in "
                                 "scipy's exact operation order, 129 contracted; the default launch keeps the exact order for "
                                 "the segmentation channel's planes only); see roofline_fp64"},
            "roofline_fp64": {"bound": "fp64_pipe", "achieved": fp64_ach, "peak": k["fp64_peak_tinstr_s"],
                              "unit": "T DP-instr/s", "frac": fp64_ach / k["fp64_peak_tinstr_s"],
                              "peak_source": "amt_fp64_probe (DMUL+DADD chains) timed in this run"},
            "kernels_ms": {"dog_axis0": k["ms_axis0"], "dog_axis1": k["ms_axis1"], "planes": k["planes"],
                           "exact_planes": k["planes"] // C if mixed else k["planes"],
                           "all_planes_exact": {"dog_axis0": k_exact["ms_axis0"], "dog_axis1": k_exact["ms_axis1"]}},
            # what binds the whole path: the float64 Gaussians need 2 * (taps of both filters) + 1 DP instructions per input
            # sample (401 in scipy's order, 269 contracted); at the measured DP issue rate that caps one GPU at this many FOV/s
            "path_fp64_ceiling": {"dp_instr_per_sample": 2 * dp_per_sample_axis + 1,
                                  "fov_per_s_per_gpu": k["fp64_peak_tinstr_s"] * 1e12 / (C * H * W * (2 * dp_per_sample_axis + 1)),
                                  "frac": (args.steps * n_fov / dev_s) / (k["fp64_peak_tinstr_s"] * 1e12 / (C * H * W * (2 * dp_per_sample_axis + 1)))},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single()
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


_REAL_STDOUT_FD: int | None = None


def capture_stdout() -> None:
    """Native libraries (NCCL's version banner, for one) write to fd 1.  Rank 0's stdout must be exactly
    one JSON line, so fd 1 points at stderr while the benchmark runs and is restored by emit()."""
    global _REAL_STDOUT_FD
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        try:
            import ctypes

            ctypes.CDLL(None).fflush(None)  # C stdio buffers of native libraries
        except Exception:
            pass
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _REAL_STDOUT_FD is not None:
        os.dup2(2, 1)  # anything printed at teardown goes to stderr again


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fovs", type=int, default=256, help="FOVs per GPU per step (config 2: 256)")
    ap.add_argument("--unique", type=int, default=8, help="distinct seeded cell layouts")
    ap.add_argument("--chunk", type=int, default=8, help="FOVs per launch wave")
    ap.add_argument("--e2e-fovs", type=int, default=256)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-alloc", default="torch", choices=["torch", "pinned", "wc"],
                    help="pinned host input buffers of the e2e leg: torch's allocator, amt_host_alloc, or amt_host_alloc write-combined")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-contracted", action="store_true",
                    help="skip the two comparison passes (every channel contracted / every channel in scipy's exact order)")
    ap.add_argument("--exact-all-channels", action="store_true",
                    help="measure FovPipelineConfig(exact_all_channels=True) as the reported configuration")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
