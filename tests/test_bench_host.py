"""Host-side logic of bench.py that runs without a GPU: the CPU-thread accounting of the reference arm (cgroup quota against
visible CPUs) and the `--impl reference` line itself (contract keys, the oracle as the thing timed)."""
import json
import os
import subprocess
import sys

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpu_threads_never_exceed_quota_or_visibility():
    threads, vis, quota = bench.cpu_threads_to_use()
    assert 1 <= threads <= vis
    if quota is not None:
        assert quota > 0 and threads <= max(1, int(quota + 0.5))


def test_cgroup_quota_parser_reads_cpu_max(tmp_path, monkeypatch):
    """cpu.max = "<quota> <period>" -> CPUs; "max" -> no limit.  The parser walks the cgroup path of /proc/self/cgroup upwards."""
    fake = tmp_path / "cpu.max"
    real_open = open

    def fake_open(path, *a, **k):
        if str(path).endswith("cpu.max"):
            return real_open(fake, *a, **k)
        if str(path).endswith("cpu.cfs_quota_us") or str(path).endswith("cpu.cfs_period_us"):
            raise OSError("no v1 hierarchy")
        return real_open(path, *a, **k)

    monkeypatch.setattr("builtins.open", fake_open)
    fake.write_text("1600000 100000\n")
    assert bench.cgroup_cpu_quota() == 16.0
    fake.write_text("max 100000\n")
    assert bench.cgroup_cpu_quota() is None
    fake.write_text("250000 100000\n")
    assert bench.cgroup_cpu_quota() == 2.5


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` times the oracle on host cores and prints ONE JSON line with the contract's keys."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--pairs", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
