"""World-size-2 run of the CB model itself on two GPUs (skipped on a one-GPU box): the row-band
split of cbinfer_b200/spatial.py (24-row input halo, NCCL send/recv, all_gather of the bands) must
reproduce the full-frame change-based model bit for bit, and stream sharding must give every rank
its own streams.  CPU/gloo twins of the host logic: tests/test_streams_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _torchrun(script_args, nproc=2, timeout=600):
    # order-fixed contraction: whole K ranges per CTA (no stream-K / cluster split-K, whose partial-sum
    # order depends on the change count of the map a rank holds), so the band result can be compared
    # bit for bit with the full frame; with the default settings the two differ by fp32 summation
    # order (~1e-7 relative) plus the threshold decisions that flips
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CBINFER_STREAMK="0", CBINFER_KSPLIT="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29531"] + script_args
    p = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("rate", [0.05, 1.0])
def test_row_band_split_on_two_gpus_is_bit_identical(rate):
    rows = _torchrun([os.path.join(REPO, "benchmarks", "split_4k.py"), "--height", "480", "--width", "640",
                      "--frames", "7", "--rate", str(rate), "--check"])
    assert rows and rows[-1]["n_gpus"] == 2
    assert rows[-1]["max_abs_diff_vs_full_frame"] == 0.0
    assert rows[-1]["ref_max_abs"] > 0.0
