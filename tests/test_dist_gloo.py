"""N > 1 host logic on CPU (gloo, world_size 2, rendezvous on 127.0.0.1): the tensor-parallel exchanges on sharded QM_CUDA tensors and
bench.py's rank plumbing for the reference arm (rank 0 alone runs and prints, the others exit 0)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import capi

ROOT = Path(__file__).resolve().parents[1]


def torchrun(nproc, port, *cmd, timeout=300):
    full = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1", "--master-port",
            str(port), *cmd]
    return subprocess.run(full, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_tensor_parallel_exchanges_world2_gloo():
    r = torchrun(2, 29621, str(ROOT / "tests" / "dist_worker.py"))
    assert r.returncode == 0 and "GLOO_TP_OK 2" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_tensor_parallel_decode_steps_world2_gloo():
    """Whole decode steps, tensor parallel over 2 ranks on the host (shard_weights + the pinned llama_forward composition + gloo collectives),
    equal the single-rank steps of the same weights: the host-side contract of BASELINE config 5."""
    r = torchrun(2, 29623, str(ROOT / "tests" / "dist_step_worker.py"))
    assert r.returncode == 0 and "GLOO_TP_STEP_OK 2" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@pytest.mark.skipif(not capi.ref_available("avx"), reason="reference AVX build (oracle/_ref) not present")
def test_bench_reference_arm_under_torchrun_prints_one_line():
    r = torchrun(2, 29622, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--model", "tiny-gqa")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0
    # what the linears-only value leaves out is measured on the host, not assumed
    ex = d["excluded_attention"]
    assert "error" not in ex and ex["attention_core_ms_per_token"] >= 0 and ex["tok_s_with_attention_at_that_ctx"] <= d["value"]
