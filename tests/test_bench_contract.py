"""
CPU: the reference arm of bench.py (`--impl reference`, the CPU path the driver times beside the
CUDA arm) runs without a GPU and prints exactly one JSON line with the keys of the bench contract.
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-sample", "4", "--height", "160", "--width", "192", "--segments", "48", "--hidden", "32",
           "--layers", "2"] + extra
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line():
    r = _run([])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"] == "images/sec graph-build+GCN trimap" and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference" = the reference's own files (oracle/_ref or /root/reference) over the shims;
    # "port" = the oracle restatement, only when those files are absent
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_port_fallback():
    r = _run([], env={"GG_CPU_PORT": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_both_arms_share_the_config_keys():
    """`config` holds the workload keys only and is produced by one function for both arms."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    a = argparse.Namespace(batch=256, height=320, width=480, segments=300, hidden=128, layers=6, nonlocal_k=4, radius=8)
    cfg = b.workload_config(a)
    assert cfg["workload"].startswith("B:") and set(cfg) == {"workload", "batch_per_gpu", "height", "width",
                                                             "n_segments", "hidden", "n_layers", "n_nonlocal",
                                                             "radius", "threshold"}
    assert b.config_letter(64, 1080, 1920) == "C" and b.config_letter(8, 2160, 3840) == "E"
    assert b.config_letter(1, 320, 480) == "A" and b.config_letter(3, 100, 100) == "custom"


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
