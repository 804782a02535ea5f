"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line, and the GPU arm
refuses to run without a device (no CPU path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--reads", "600", "--read-len", "4000", "--genomes", "2", "--no-stages"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Gbp/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 2 and j["n_gpus"] == 1 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["k"] == 4 and j["config"]["minimizer_size"] == 15 and "workload" in j["config"]
    assert j["check"]["n_solid"] > 0


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--reads", "100"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout) or "CUDA" in (out.stderr + out.stdout)


def test_gpu_arm_runs_on_the_emulator(tmp_path):
    """bench.py's GPU arm, end to end, on the CPU emulator of the C-ABI library with the test-only torch stand-in
    (tests/mock_torch): autotune, device-resident leg, multi-k extra, host-buffer e2e leg with its table checks, CPU
    baseline + fingerprint comparison, and the one JSON line with every contract key.  The numbers mean nothing
    (emulated kernels); the point is that the judged run cannot die on a Python error."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _emu
    lib = _emu.build_emulated_library(str(tmp_path))
    env = dict(os.environ, MDBG_EMU_LIB=lib)
    for extra in (["--extras", "cfg4,cfg5", "--extra-scale", "0.00002"], ["--workload", "cfg3"]):
        run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_on_emulator.py"), "--reads", "240",
                              "--read-len", "3000", "--genomes", "2", "--steps", "2", "--warmup", "3", "--e2e-batch", "100",
                              "--multi-k", "6"] + extra, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-3000:]
        lines = [ln for ln in run.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1, run.stdout[-2000:]
        d = json.loads(lines[0])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert key in d, key
        assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
        assert d["e2e"]["table_checks"]["steps_differing_from_device_leg"] == 0 and d["e2e"]["table_checks"]["steps_checked"] == 4
        assert d["check"]["occurrences_conserved"] and d["check"]["device_steps_same_checksum"]
        assert all(d["sketch_autotune"]["identical"]) and d["multi_k"]["same_tables_both_sweeps"]
        assert d["edges"]["n_edges"] > 0 and d["edges"]["n_nodes"] == d["check"]["n_solid_total"]
        if "--extras" in extra:
            assert d["unitigs"]["n_nodes"] == d["edges"]["n_nodes"] and d["unitigs"]["n_unitigs"] > 0, d["unitigs"]
            assert d["unitigs"].get("cpu_reference_on_sample", {}).get("identical_records", d["cpu_baseline"]["kind"] == "port"), d["unitigs"]
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and "identical" in d["cpu_baseline"]["sample"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert d["ascii_resident"]["same_table_as_packed_leg"] and d["ascii_resident"]["value"] > 0
        if "--extras" in extra:                      # the other BASELINE configs run through the same step, scaled down
            x4, x5 = d["extras"]["cfg4"], d["extras"]["cfg5"]
            assert "error" not in x4 and "error" not in x5, (x4, x5)
            assert x4["scaling"] == "strong" and x4["k_last"] == 6 and len(x4["n_entries_per_k_total"]) == 3
            assert x5["config"]["samples"] == 3 and x5["n_minimizers_total"] > 0 and x5["value"] > 0
