"""bench.py's driver contract on the CPU arm (no GPU needed): `--impl reference` prints exactly ONE JSON line on stdout —
even when a library writes to the C-level stdout behind Python's back — with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1", "--cpu-batch", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training images/sec (G+D step)" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"] == "train_step_64x64_b4096_per_gpu"
    # the unmodified reference from oracle/_ref when that copy exists (built by oracle/make_ref.py), else the oracle port
    ref_copy = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json"))
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_copy else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_stdout_is_claimed_before_libraries_can_write_to_it(tmp_path):
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('stray print'); "
            "print(json.dumps({'ok': 1}), file=bench.RESULT_OUT, flush=True)" % ROOT)
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    assert proc.stdout.strip() == '{"ok": 1}'
    assert "NCCL version banner" in proc.stderr and "stray print" in proc.stderr
