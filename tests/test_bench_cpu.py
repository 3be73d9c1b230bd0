"""bench.py's clock sampler (host logic, no GPU): rows are attributed to the timed region by arrival time, a region shorter
than one polling period falls back to the warm-up's last rows and says so, throttle reasons are picked up."""
import importlib.util
import os
import stat
import time

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("fj_bench", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _fake_smi(tmp_path, row, period_s):
    """A stand-in for `nvidia-smi --query-gpu=... -lms N`: prints `row` every `period_s` seconds until it is terminated."""
    exe = tmp_path / "nvidia-smi"
    exe.write_text("#!/bin/sh\nwhile true; do echo '%s'; sleep %s; done\n" % (row, period_s))
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    return str(tmp_path)


OK_ROW = "0, 1965, 1965, 512.30, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active"
HOT_ROW = "0, 1410, 1965, 990.00, 0x0000000000000040, Not Active, Active, Not Active, Active"


def test_rows_of_the_timed_region_only(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, OK_ROW, 0.05) + os.pathsep + os.environ["PATH"])
    s = _bench().ClockSampler(0)
    s.start()
    time.sleep(0.3)                      # "warm-up": rows arrive, none of them may be reported
    before = len(s.rows)
    s.mark()
    time.sleep(0.4)
    out = s.stop()
    assert before >= 2
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    assert 2 <= out["samples"] <= len(s.rows) - before + 1
    assert "note" not in out


def test_region_shorter_than_the_period_falls_back_to_the_warmup(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, OK_ROW, 0.2) + os.pathsep + os.environ["PATH"])
    s = _bench().ClockSampler(0)
    s.start()
    time.sleep(0.5)
    s.mark()
    out = s.stop()                       # no time for a row in between
    assert out["samples"] >= 1 and out["sm_mhz"] == 1965.0
    assert "warm-up" in out["note"]


def test_throttle_reasons_are_reported(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, HOT_ROW, 0.05) + os.pathsep + os.environ["PATH"])
    s = _bench().ClockSampler(0)
    s.start()
    s.mark()
    time.sleep(0.3)
    out = s.stop()
    assert out["sm_mhz"] == 1410.0
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]


def test_period_comes_from_the_environment(monkeypatch):
    monkeypatch.setenv("FJ_CLOCK_LMS", "250")
    assert _bench().ClockSampler(0).period_ms == 250
    monkeypatch.delenv("FJ_CLOCK_LMS")
    assert _bench().ClockSampler(0).period_ms == 1000


def test_wait_first_row_returns_once_nvml_is_up(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, OK_ROW, 0.05) + os.pathsep + os.environ["PATH"])
    s = _bench().ClockSampler(0, period_ms=700)
    assert s.period_ms == 700
    s.start()
    t0 = time.perf_counter()
    s.wait_first_row(timeout_s=5.0)
    assert len(s.rows) >= 1 and time.perf_counter() - t0 < 4.0
    s.mark()
    out = s.stop()
    assert out["period_ms"] == 700
