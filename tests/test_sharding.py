"""N > 1 host logic on CPU: two gloo ranks shard a plate of FOVs, each produces its per-FOV
tables, rank 0 gathers them in FOV order; the benchmark's max-over-ranks timing reduce."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arcadia_microscopy_tools_b200.sharding import gather_fov_results, max_over_ranks, shard_indices


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_fov: int, out_path: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_indices(n_fov, rank, world)
        # stand-in for the per-FOV cell table: deterministic function of the global index
        local = {int(i): {"area": np.arange(int(i) % 5 + 1, dtype=np.float64) * (i + 1), "rank": rank} for i in mine}
        merged = gather_fov_results(local, n_fov, dist)
        slowest = max_over_ranks(1.0 + rank, dist)
        assert slowest == float(world)
        if rank == 0:
            assert merged is not None and len(merged) == n_fov
            for i, item in enumerate(merged):
                assert item["rank"] == i % world
                assert np.array_equal(item["area"], np.arange(i % 5 + 1, dtype=np.float64) * (i + 1))
            with open(out_path, "w") as fh:
                fh.write("ok")
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_shard_indices_cover_everything_once():
    for n, g in [(0, 2), (1, 2), (9, 2), (3456, 8), (7, 8)]:
        parts = [shard_indices(n, r, g) for r in range(g)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_single_process_gather_and_checks():
    local = {i: i * i for i in range(5)}
    assert gather_fov_results(local, 5) == [0, 1, 4, 9, 16]
    with pytest.raises(ValueError, match="not processed"):
        gather_fov_results({0: 1}, 2)
    assert max_over_ranks(0.25) == 0.25


def test_two_gloo_ranks_shard_and_gather(tmp_path):
    out = tmp_path / "done"
    mp.spawn(_worker, args=(2, _free_port(), 11, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"



def _fake_process(fovs, givens):
    """Stand-in for the executor: 'count' = first pixel of channel 0, a one-column table carrying the FOV's sum."""
    n = fovs.shape[0]
    out = {"thresholds": np.zeros(n), "status": np.zeros(n, np.int32)}
    for which in ("thr", "given"):
        out[f"counts_{which}"] = fovs[:, 0, 0, 0].astype(np.int32)
        tab = np.zeros((n, 2, 8))
        tab[:, 0, :] = fovs.reshape(n, -1).sum(axis=1)[:, None]
        out[f"tables_{which}"] = tab
    return out


def _plate_source(i):
    fov = np.full((2, 4, 4), i % 7 + 1, dtype=np.uint16)
    return fov, np.zeros((4, 4), np.int32)


def _plate_worker(rank: int, world: int, port: int, n_fov: int, out_path: str) -> None:
    from arcadia_microscopy_tools_b200.batch import FovPipelineConfig, run_plate

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = FovPipelineConfig(n_channels=2, height=4, width=4, max_labels=8)
        merged = run_plate(_plate_source, n_fov, cfg, dist=dist, batch_fovs=3, process_batch=_fake_process)
        if rank == 0:
            assert len(merged) == n_fov
            for i, rec in enumerate(merged):
                assert rec["rank"] == i % world and rec["count_thr"] == i % 7 + 1
                assert rec["table_thr"].shape == (2, i % 7 + 1) and rec["table_thr"][0, 0] == 32 * (i % 7 + 1)
            with open(out_path, "w") as fh:
                fh.write("ok")
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_run_plate_two_gloo_ranks(tmp_path):
    out = tmp_path / "plate_ok"
    mp.spawn(_plate_worker, args=(2, _free_port(), 17, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
