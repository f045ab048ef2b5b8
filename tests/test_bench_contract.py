"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`)
times the unmodified reference (oracle/_ref/lmp_ref) on the host cores and prints ONE JSON line
with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if not (ROOT / "oracle" / "_ref" / "lmp_ref").exists():
        pytest.skip("oracle/_ref/lmp_ref not built (python oracle/build_ref.py)")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload",
                        "lj32k", "--steps", "3", "--warmup", "3"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "atom-timesteps/s" and d["unit"] == "atom-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "lj32k", "--steps", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0          # no CPU fallback: the product path needs the CUDA device
    assert not r.stdout.strip()
