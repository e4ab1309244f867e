"""CPU: host-side pieces of bench.py that do not need a GPU -- the clock sampler (time-stamped nvidia-smi lines, load-window
filter, throttle-reason parsing) against a fake `nvidia-smi`, and the shape of the reference-arm JSON line."""
import os
import stat
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fake_smi(tmp_path, line):
    p = tmp_path / "nvidia-smi"
    p.write_text("#!/bin/bash\nwhile true; do echo '%s'; sleep 0.02; done\n" % line)
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(tmp_path)


def test_clock_sampler_window_and_reasons(tmp_path, monkeypatch):
    import bench
    d = _fake_smi(tmp_path, "0, 1965, 1965, 410.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active")
    monkeypatch.setenv("PATH", d + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.15)
    t0 = time.time()
    time.sleep(0.25)
    t1 = time.time()
    assert s.count_between(t0, t1) >= 3
    assert s.count_between(t1 + 10, t1 + 11) == 0
    out = s.stop(t0, t1)
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
    assert 3 <= out["samples"] <= 40 and 200 <= out["window_ms"] <= 400


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    import bench
    monkeypatch.setenv("PATH", str(tmp_path))              # no nvidia-smi anywhere
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop(0.0, 1.0)
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]
