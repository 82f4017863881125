"""The measurement contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`, the unmodified
reference's OpenMP build on the host cores) prints ONE JSON line with the arm's metric, unit and config, a cpu_baseline that
describes the run and an e2e block without device copies; under torchrun rank 0 alone runs it and the other ranks exit 0."""
import json
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "omp" / "librefshim.so"


def check_line(out: str, n_gpus: int):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["n_gpus"] == n_gpus
    assert "4K IYUV DCT-50 compress+decompress" in d["metric"] and d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["config"]["width"] == 3840 and d["config"]["height"] == 2160 and d["config"]["quality"] == 50 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and "sample" in cb
    assert len(cb["omp_settings_tried"]) == 2 and max(cb["omp_settings_tried"].values()) == d["value"]  # both nestings run, the better reported
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


@pytest.mark.timeout(600)
def test_reference_arm_line():
    if not REF.exists():
        pytest.skip("oracle/_ref (the unmodified reference) is not built")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-frames", "1"],
                       capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    check_line(r.stdout, 1)


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    if not REF.exists():
        pytest.skip("oracle/_ref (the unmodified reference) is not built")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--cpu-frames", "1"], capture_output=True, text=True, timeout=580, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    check_line(r.stdout, 2)  # one line for the whole job: the other rank printed nothing


def test_product_arm_has_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and not [l for l in r.stdout.splitlines() if l.startswith("{")]
