"""The stored bench lines (profiles/r0N_bench_*.json, written by bench.py on the GPU box) carry every key of the bench
contract; a guard against a bench.py edit that silently drops one.  Round 1 lines: config C2 as a batch, one sequence
per GPU (weak scaling).  Round 2 lines: config C5, the same 4096 pairs split over the GPUs (strong scaling), with the
other configurations as named sub-results."""
import glob
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config"]


def lines():
    out = []
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[12]_bench_*.json"))):
        for ln in open(f):
            if ln.startswith("{"):
                out.append((os.path.basename(f), json.loads(ln)))
    return out


@pytest.mark.parametrize("name,d", lines())
def test_bench_line_has_the_contract_keys(name, d):
    for k in BASE:
        assert k in d, (name, k)
    assert d["metric"] == "ICP scan-pairs/s" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) if d.get("cpu_baseline") else True
    if d.get("impl") == "reference":
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
        assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
        return
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if name.startswith("r01"):
        pairs = int(re.search(r"C2: (\d+)-pair", d["config"]["workload"]).group(1))   # per GPU (weak scaling)
        assert abs(d["value"] - d["n_gpus"] * pairs / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6  # whole-job aggregate
    else:
        pairs = int(re.search(r"C5: batched ICP of (\d+) independent", d["config"]["workload"]).group(1))
        assert d["scaling"] == "strong" and d["workload_stats"]["pairs"] == pairs
        assert abs(d["value"] - pairs / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6   # the job is always the same pairs
        assert d["workload_stats"]["status_ok_frac"] == 1.0
        assert 0 < d["e2e"]["h2d_ceiling"]["e2e_frac_of_ceiling"] <= 1.05
        c4 = d["c4_loop_closure"]
        assert c4["n_gpus"] == d["n_gpus"] and c4["ms_per_detect"] > 0 and c4["top10_that_are_true_revisits"] >= 3
        if d["n_gpus"] == 1:
            assert d["c2_streaming"]["ms_per_frame_mean"] > 0 and d["c2_batch"]["pairs_per_s"] > 0
            assert d["c3_knn_normals"]["knn_plus_normals_queries_per_s"] > 0
            assert d["roofline"]["issue"] is None or 0 < d["roofline"]["issue"]["frac"] < 1
    if d["n_gpus"] == 1:
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["cpu_baseline"]["cores"] >= 1
