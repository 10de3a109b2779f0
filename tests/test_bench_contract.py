"""CPU: the bench lines committed under profiles/r2/ (printed by bench.py on B200s) carry every key of the bench contract,
and the reference arm -- which needs no GPU -- prints a contract-shaped line here."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}


@pytest.mark.parametrize("name", ["final2_c2.json", "final2_c3.json", "final2_c4.json", "final2_g8_c4.json", "head_c5.json", "head_g8_c5.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = json.loads((ROOT / "profiles" / "r2" / name).read_text())
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["metric"] == "converged_nmpc_solves_per_sec" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].split("_")[0] in json.dumps(base).lower()                  # BASELINE.json's metric: converged NMPC solves/s
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 8 and "workload" in d["config"] and "l2" in d["config"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["steps"] / 1e3 - d["wall_s"]) < 0.2 * d["wall_s"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1 and d["cpu_baseline"] is not None:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    # status census: every solve of the timed region is accounted for, converged ones are what `value` counts
    assert abs(d["status_hist"][0] / sum(d["status_hist"]) - d["converged_fraction"]) < 1e-9


def test_reference_arm_line_here():
    """bench.py --impl reference times the CPU restatement (the oracle) and needs no GPU: run it small and check the line."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "64"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "converged_nmpc_solves_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
