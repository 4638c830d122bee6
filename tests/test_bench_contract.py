"""bench.py's CPU arm (`--impl reference`) and JSON contract, on tiny workloads (no GPU needed): one JSON line with
the keys the driver reads, rank != 0 prints nothing, the memory guard of the CPU arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-procs", "1"] + args, capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


@pytest.mark.parametrize("workload,metric", [("poisson", "poisson3d_gmg_cg_mdof_per_s"),
                                             ("convdiff", "convdiff3d_bicgstab_gmg_gs_mdof_per_s"),
                                             ("elasticity", "elasticity3d_gmg_cg_mdof_per_s")])
def test_reference_arm_prints_one_contract_line(workload, metric):
    lines = _run(["--workload", workload, "--refs", "2"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "MDoF/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    assert _run(["--refs", "2", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_cpu_arm_is_bounded_by_host_memory():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.replica_bytes(7) < bench.replica_bytes(8) and bench.replica_bytes(6, "elasticity") > bench.replica_bytes(6)
    assert 1 <= bench.bounded_procs(64, 9, "elasticity") <= 64           # a 513^3 x 3 hierarchy never gets 64 replicas here
    assert bench.bounded_procs(1, 2) == 1


def test_reference_arm_strong_scaling_grid_and_order():
    """--scaling strong: the CPU arm solves the same 2x2x2-cell grid (here 9^3 at numRefs 2) and says so; --order hier is
    passed through to the generator (same iteration count, another numbering)."""
    d = json.loads(_run(["--refs", "2", "--scaling", "strong"])[0])
    assert d["scaling"] == "strong" and "729 DoF" in d["cpu_baseline"]["sample"]
    h = json.loads(_run(["--refs", "2", "--order", "hier"])[0])
    l = json.loads(_run(["--refs", "2"])[0])
    assert h["config"]["iterations"] == l["config"]["iterations"]


def test_gpu_arm_json_fields_are_documented_in_the_source():
    """The GPU arm cannot run here; pin the keys it emits by reading the source: the driver's contract keys plus the
    roofline objects of DESIGN.md §4."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"roofline"', '"roofline_plain"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"', '"h2d_bytes_per_step"',
                '"frac_vs_spec"', '"crs_bytes_per_launch"', '"traffic"', '"history"', '"scaling": "strong" if strong else "weak"'):
        assert key in src, key
