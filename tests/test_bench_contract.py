"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line, the clock sampler's
parsing of nvidia-smi rows, and that the product arm fails loudly (no CPU fallback) when there is no device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_one_contract_line(built):
    from oracle import oracle_py as O

    if not O.have_reference():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines  # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "composited_mblocks_per_s" and d["unit"] == "Mblocks/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_clock_sampler_parses_nvidia_smi_rows():
    b = _load_bench()
    s = b.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.lines = [(10.0, "0, 1965, 1965, 512.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active"),
               (10.1, "0, 1950, 1965, 530.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active"),
               (10.2, "garbage"), (10.3, "0, 1965, 1965, 500.0, 0x0, Not Active, Not Active, Not Active, Not Active")]
    c = s.stop(9.9, 10.4)
    assert c["samples"] == 3 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]


def test_product_arm_fails_loudly_without_a_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert not out.stdout.strip()  # no result line: nothing was measured
