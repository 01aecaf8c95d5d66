"""bench.py's reference arm runs on the CPU (the oracle port of the reference's path) and prints the contract's JSON line;
the B200 arm refuses to run without a GPU instead of falling back."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["metric"] == "routed_multi_lora_whisper_fwd_clips_per_sec" and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert set(line["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
