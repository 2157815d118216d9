"""GPU, 2 ranks under torchrun (skipped with fewer than 2 devices): the data-parallel fine-tuning step of
ModelFinetuner.train -- the path's ONE collective -- against the same step on one rank with the doubled batch.

Reference semantics (finetune_trainer.py:939-957, :1007-1023; sequence_tagger_model.py:2506): the loss is a per-batch
mean, so the mean of the rank gradients is the gradient of the world*B batch; clipping uses the post-reduce norm; every
rank applies the same update (replicas stay identical)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(cmd, env=None):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("payload,overlap", [("bf16", "0"), ("fp32", "0"), ("bf16", "1")])
def test_two_ranks_equal_one_rank_with_the_doubled_batch(tmp_path, payload, overlap):
    """payload bf16 = packed buffers + sparse embedding rows + row-skipping optimizer; fp32 = dense in-place all-reduce;
    overlap 1 = the chunked backward with the exchange of finished layer chunks started under the next one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    worker = os.path.join(HERE, "ddp_worker.py")
    port = str(29700 + os.getpid() % 200)
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
          "--master-port", port, worker, str(tmp_path)], env={"KBNER_GRAD_COMM": payload, "KBNER_OVERLAP_ALLREDUCE": overlap})
    _run([sys.executable, worker, str(tmp_path), "--single", "2"], env={"CUDA_VISIBLE_DEVICES": "0"})
    r0 = torch.load(tmp_path / "params-ddp-rank0.pt")
    r1 = torch.load(tmp_path / "params-ddp-rank1.pt")
    one = torch.load(tmp_path / "params-single-rank0.pt")
    lr = 1e-3
    worst = 0.0
    for k in one:
        assert torch.equal(r0[k], r1[k]), "replicas diverged at %s" % k      # same reduced gradient, same update
        # AdamW's first step moves every element by ~lr * sign(g): compare the UPDATES, in units of lr.  fp32 payload:
        # only the summation order differs; bf16 payload: one bf16 rounding per addend (sign flips only where |g| ~ 0)
        diff = (r0[k] - one[k]).abs() / lr
        frac_off = float((diff > 0.25).float().mean())
        worst = max(worst, frac_off)
        assert frac_off < (0.02 if payload == "bf16" else 0.01), (k, frac_off)
        assert float(diff.mean()) < 0.05, (k, float(diff.mean()))
    print("DDP vs single (%s payload, overlap %s): worst fraction of elements whose update differs by > lr/4: %.4f"
          % (payload, overlap, worst))
