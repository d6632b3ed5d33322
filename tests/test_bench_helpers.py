"""CPU checks of bench.py's host-side helpers (no GPU): the clock sampler degrades to "no samples" without
NVML / nvidia-smi, pause/resume are harmless in every mode, and the reference arm prints one JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_clock_sampler_without_gpu_tools():
    import bench
    s = bench.ClockSampler(0, period_s=0.01)
    s.start()
    s.pause()
    s.resume()
    out = s.stop()
    # no GPU in the CPU container: either nothing could be sampled (None) or a well-formed record
    assert out is None or {"sm_mhz", "sm_max_mhz", "reasons", "samples"} <= set(out)


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
