"""Host-side pieces of bench.py that run without a GPU: the clock sampler's helper-process protocol (with a stand-in for
NVML), the workload table (every BASELINE.json config selectable, identical `config` object in both arms) and the
algorithmic-flop count the step-level roofline divides by."""
import json
import os
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture()
def bench(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    import bench as B
    yield B
    B.select_workload("cfg2")


def test_clock_sampler_helper_process_samples_only_between_begin_and_stop(bench, monkeypatch):
    monkeypatch.setenv("DVAE_FAKE_NVML", "1")
    c = bench.ClockSampler(0, period_s=0.002)
    assert c.child is not None and c.source == "nvml helper process"
    time.sleep(0.05)                     # before begin(): nothing may be recorded
    c.begin()
    time.sleep(0.05)
    c.sample_now()                       # a no-op now: the launch loop never calls NVML
    out = c.stop()
    assert 5 <= out["samples"] <= 40, out
    assert out["sm_mhz"] == out["sm_max_mhz"] == 1965.0 and out["reasons"] == [] and out["source"] == "nvml helper process"
    assert c.child.poll() is not None    # the helper has exited


def test_clock_sampler_reports_throttle_reasons(bench):
    c = bench.ClockSampler.__new__(bench.ClockSampler)
    c.samples, c.live, c._stop, c.h, c.child, c.source, c.max_mhz = [(1500.0, 0x4 | 0x40, 700.0), (1965.0, 0, 500.0)], False, False, object(), None, "test", 1965.0
    out = c.stop()
    assert set(out["reasons"]) == {"sw_power_cap", "hw_thermal_slowdown"} and out["sm_min_mhz"] == 1500.0 and out["power_w_max"] == 700.0


def test_clock_sampler_without_nvml_degrades_to_null_clocks(bench, monkeypatch):
    monkeypatch.delenv("DVAE_FAKE_NVML", raising=False)
    c = bench.ClockSampler(0)            # no GPU / driver in the CPU container: neither the helper nor the thread can sample
    c.begin()
    out = c.stop()
    if out["samples"] == 0:
        assert out["sm_mhz"] is None and out["reasons"] == []


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_every_baseline_config_is_a_selectable_workload(bench, name):
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    bench.select_workload(name)
    cfg = bench.workload_config(1)
    assert cfg["workload"].startswith(name) and cfg["global_batch"] == bench.BATCH and cfg["seq_len"] == bench.SEQ_T
    assert bench.workload_config(8)["global_batch"] == 8 * bench.BATCH and bench.workload_config(8)["parallelism"] == "dp8"
    assert "model" not in cfg            # the tier's contract: config names the workload, no model keys


def test_algorithmic_flops_match_the_survey_accounting(bench):
    """SURVEY.md 8a / DESIGN.md 6: 102.6 GFLOP algorithmic per cfg-2 train step, forward + backward = 3 x forward, the
    vocabulary projection the largest single class."""
    bench.select_workload("cfg2")
    f = bench.algorithmic_flops(bench.BATCH, bench.SEQ_T)
    n = (bench.SEQ_T - 1) * bench.BATCH
    assert f["vocab"] == 3.0 * 2 * n * bench.CFG2["hidden_dim"] * bench.VOCAB
    assert abs(f["total"] - 102.64e9) < 0.05e9
    assert f["total"] == f["vocab"] + f["lstm_input_projections"] + f["recurrence"] + f["heads"]
    assert f["vocab"] > max(f["lstm_input_projections"], f["recurrence"], f["heads"])
    bench.select_workload("cfg4")
    f4 = bench.algorithmic_flops(bench.BATCH, bench.SEQ_T)
    assert f4["vocab"] == 3.0 * 2 * (bench.SEQ_T - 1) * bench.BATCH * 1024 * 50000
