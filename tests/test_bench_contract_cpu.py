"""bench.py's reference arm (the CPU leg the driver launches as `bench.py --impl reference`) on a one-utterance sample:
one JSON line with the contract's keys, no GPU needed; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--cpu-sample", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run_bench()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("audio-sec/sec") and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "utterances" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2:")
    # the line reports the steps it RAN (the arm is bounded to 10 timed steps) and keeps the request alongside
    assert d["steps"] == 1 and d["steps_requested"] == 1 and "1 timed runs" in cb["sample"]
    # both arms print the same `config` dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    ours = bench.shared_config(argparse.Namespace(workload="c2"), bench.WORKLOADS["c2"], 1)
    assert d["config"] == ours


def test_reference_arm_is_silent_on_other_ranks():
    assert run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
