"""The JSON line ``bench.py`` prints is a contract with the driver; the committed lines of the last measurement
pass (``profiles/r02_bench_*.json``) must carry every key it names, with consistent values."""

from __future__ import annotations

import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _line(name: str) -> dict:
    path = PROFILES / name
    if not path.exists():
        pytest.skip(f"{name} not committed")
    return json.loads(path.read_text())


@pytest.mark.parametrize("name,n_gpus", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n8.json", 8)])
def test_b200_line_has_the_contract_keys(name, n_gpus):
    d = _line(name)
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert d["n_gpus"] == n_gpus and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0
    # value = samples per second of the timed steps (256 FOVs of 4 x 2048 x 2048 per GPU and step)
    samples = n_gpus * d["steps"] * d["config"]["fovs_per_gpu"] * 4 * 2048 * 2048
    assert d["value"] == pytest.approx(samples / (d["ms_per_step"] * d["steps"] * 1e-3) / 1e6, rel=1e-6)
    assert d["config"]["uses_tensor_cores"] is True and d["config"]["decision_exact"] is True and d["config"]["exact_planes"] == 0
    e2e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "copy_only", "frac_of_copy_ceiling", "uint16_masks"} <= set(e2e)
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < d["value"]
    assert e2e["fovs_per_step"] == 256  # the same batch at every N
    # int64 label masks in host memory: 4 x 2 bytes of pixels + 8 bytes of label per FOV-pixel read by the call; when host
    # threads narrow the masks to uint16 inside the call, 2 bytes of label cross PCIe and the int64 route is reported beside it
    px = e2e["fovs_per_step"] * 2048 * 2048
    if "masks_sent_as_runs" in e2e:
        # label masks cross PCIe as per-row runs where the host threads keep up (the executor balances per chunk); the
        # plain routes are reported beside it
        n_runs, n_plain = e2e["masks_sent_as_runs"], e2e["masks_sent_plain"]
        assert n_runs + n_plain == e2e["fovs_per_step"] and e2e["host_bytes_read_per_step"] == px * (4 * 2 + 8)
        lo = px * 4 * 2 + n_plain * 2048 * 2048 * 8
        assert lo < e2e["h2d_bytes_per_step"] <= lo + n_runs * 2048 * 2048 * 2
        assert e2e["int64_over_pcie"]["h2d_bytes_per_step"] == px * (4 * 2 + 8) and 0 < e2e["int64_over_pcie"]["value"] < e2e["value"]
        assert e2e["int64_narrowed_to_uint16"]["h2d_bytes_per_step"] == px * (4 * 2 + 2)
    elif "host_bytes_read_per_step" in e2e:
        assert e2e["host_bytes_read_per_step"] == px * (4 * 2 + 8) and e2e["h2d_bytes_per_step"] == px * (4 * 2 + 2)
        assert e2e["int64_over_pcie"]["h2d_bytes_per_step"] == px * (4 * 2 + 8) and 0 < e2e["int64_over_pcie"]["value"] < e2e["value"]
    else:
        assert e2e["h2d_bytes_per_step"] == px * (4 * 2 + 8)
    assert 0.5 < e2e["frac_of_copy_ceiling"] <= 1.05 and e2e["uint16_masks"]["value"] > 0.98 * e2e["value"]
    clocks = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clocks) and clocks["samples"] >= 10
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
    roof = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof) and roof["bound"] in ("hbm", "tensor")
    assert roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"], rel=1e-9) and 0 < roof["frac"] < 1
    assert "tcg_axis1" in roof["kernel"] and roof["traffic"] == pytest.approx(roof["algorithmic_bytes_per_launch"], rel=0.1)
    stages = d["stages"]
    assert {"W_pre", "W_seg", "W_quant", "stage_ms_per_chunk"} <= set(stages)
    assert d["chunk"]["ms"] > 0 and d["chunk"]["ms_budget_for_60pct_of_hbm"] == pytest.approx(1.5357 * d["chunk"]["fovs"] / 8, rel=1e-3)
    for mode in d["other_modes"].values():  # every arithmetic mode decides the same labels as the default one
        assert mode["thresholds_counts_and_tables_bit_identical_to_default"] is True
    if n_gpus == 1:
        cpu = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(cpu) and cpu["kind"] in ("reference", "port")
        assert d["parity_checked"]["ok"] is True and d["parity_checked"]["fovs"] == 3
        assert d["parity_checked"]["max_plane_abs_err_other_channels"] <= 1e-8
    else:
        plate = d["plate"]
        assert plate["fovs_gathered"] == n_gpus * 256 and "verified" in plate["order"] and plate["fovs_with_status"] == 0


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(d) and d["impl"] == "reference"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
