"""bench.py's reference arm and bookkeeping, without a GPU: the CPU arm runs the reference's own host-compiled
code (or the oracle port) on a bounded sample, prints the contract's JSON line, describes the workload with exactly
the keys the GPU arm uses (the driver compares the two `config` objects), and never maps the product library."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_line_and_config(built):
    env = dict(os.environ, ATX_BENCH_CPU_SECONDS="0.5")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpaths/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"] == {"value": line["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["product_library_loaded"] is False
    sys.path.insert(0, str(ROOT))
    import bench
    want = bench.config_dict("c2", 1, 1024, 3, 1)            # what the GPU arm prints for the same workload
    assert line["config"] == want
    assert set(bench.STRONG) <= set(bench.WORKLOADS) and set(bench.SCENE_OF) == set(bench.WORKLOADS)


def test_rank_other_than_zero_does_no_work(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                          text=True, timeout=120, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""
