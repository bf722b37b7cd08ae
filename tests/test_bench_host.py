"""Host-side pieces of bench.py (no GPU): workload description per BASELINE.json config, peak lookup, parsing of the
nvidia-smi clock samples that end up in the JSON line's "clocks" object."""
import importlib
import sys

import pytest


@pytest.fixture()
def bench(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    import bench as b
    return importlib.reload(b)


def test_default_workload_is_baseline_config_2(bench, monkeypatch):
    a = bench.parse()
    bench.apply_workload(a)
    cfg = bench.workload_config(a)
    assert (a.batch, a.micro_batch, a.N, a.snr, a.steps, a.warmup) == (16, 16, 30, 0.5, 3, 3)
    assert "VoiceBank" in cfg["workload"] and "N=30" in cfg["workload"] and cfg["baseline_config"] == 2
    assert cfg["global_batch"] == 16 and "no data-path collective" in cfg["parallelism"] and "L2" in cfg["l2"]
    assert "model" not in cfg                                   # this tier's config names a workload, not a model
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "3"])
    a = bench.parse()
    bench.apply_workload(a)
    assert (a.batch, a.N, bench.SR, bench.GFLOP_PER_FORWARD, a.no_cpu_baseline) == (8, 30, 48000, 3187.6, True)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "4", "--gpus", "2"])
    a = bench.parse()
    bench.apply_workload(a)
    cfg = bench.workload_config(a)
    assert (a.batch, a.N, a.snr) == (32, 50, 0.33) and cfg["global_batch"] == 64 and "100 network evaluations" in cfg["workload"]


def test_peaks_come_from_the_driver_file_or_the_stated_fallback(bench):
    p = bench.peaks()
    assert p["source"].startswith("MEASURED_PEAKS.json") or p["source"].startswith("fallback")
    assert 3000 < p["hbm_gbs"] < 9000 and 800 < p["tflops_sustained"] <= p["tflops_burst"] < 2500


def test_clock_samples_are_reduced_to_median_max_and_reasons(bench):
    cs = bench.ClockSampler(0)
    cs.rows = [
        ["0", "1755", "1965", "980.1", "Not Active", "Not Active", "Not Active", "Active"],
        ["0", "1650", "1965", "995.0", "Not Active", "Not Active", "Not Active", "Active"],
        ["0", "1800", "1965", "700.2", "Not Active", "Not Active", "Not Active", "Not Active"],
        ["garbage"],
        ["0", "[N/A]", "1965", "1.0", "Not Active", "Not Active", "Not Active", "Not Active"],
    ]
    out = cs.stop()
    assert out == {"sm_mhz": 1755.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    cs.rows = [["0", "600", "1965", "100", "Active", "Active", "Not Active", "Not Active"]]
    assert cs.stop()["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]
    assert bench.ClockSampler(0).stop() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_reference_arm_runs_on_rank_0_only():
    """Under torchrun (N > 1) only rank 0 times the CPU reference; the other ranks exit 0 without work or output."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tools_are_importable_python_and_shell_scripts_parse():
    """The round-2 tools were written without a GPU: at least their syntax is checked here (py_compile, bash -n)."""
    import glob
    import os
    import py_compile
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in sorted(glob.glob(os.path.join(root, "tools", "*.py"))):
        py_compile.compile(f, doraise=True)
    for f in sorted(glob.glob(os.path.join(root, "tools", "*.sh"))):
        r = subprocess.run(["bash", "-n", f], capture_output=True, text=True)
        assert r.returncode == 0, f"{f}: {r.stderr}"


def test_conv_traffic_capture_is_tied_to_the_kernel_sources(tmp_path, capsys):
    """roofline.traffic comes from ncu counters of the CURRENT build: tools/make_conv_traffic.py stores a digest of the
    dominant kernel's sources next to the counters, bench.py drops a capture whose digest differs (stale file -> null)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import make_conv_traffic as m
    csvf = tmp_path / "t.csv"
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    row = lambda name, val: f'"0","1","python","h","conv_tc6_kernel<2, 6>(...)","1","7","(448, 1, 1)","(148, 1, 1)","0","10.0","s","{name}","byte","{val}"'
    csvf.write_text("==PROF== noise\n" + hdr + "\n" + row("dram__bytes_read.sum", "537,675,008") + "\n" +
                    row("dram__bytes_write.sum", "499070464") + "\n" + row("gpu__time_duration.sum", "545760") + "\n")
    m.main(str(csvf))
    d = json.loads(capsys.readouterr().out)
    assert d["traffic_bytes_per_launch"] == 537675008 + 499070464 and d["algorithmic_bytes_per_launch"] == 2 ** 30
    assert d["source_digest"] == m.source_digest() and len(d["source_digest"]) == 16 and 0.9 < d["ratio"] < 1.0
