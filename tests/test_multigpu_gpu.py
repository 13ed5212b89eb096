"""SURVEY.md §4 tier 6 on real GPUs: the N-rank data-parallel step (SyncBN forward + backward through the peer-memory
kernel AND through NCCL, batchwise global Dice, GradReducer) equals the single-device step on the concatenated batch.
Needs >= 2 visible GPUs (torchrun, 127.0.0.1 rendezvous); the driver's single-GPU test box runs the world-size-1 leg
only, `bench.py --gpus N` runs the same check (`selfcheck.n_rank_parity`) before it times anything."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_world_size_1_leg_runs_the_same_code():
    from medsegpretrainimagenet_b200.selfcheck import exchange_parity, n_rank_parity
    ex = exchange_parity(None)
    assert ex["world"] == 1 and ex["ok"], ex
    res = n_rank_parity(None)
    assert res["world"] == 1 and res["ok"], res
    assert res["counters_exchange_bit_exact"] and res["loss_rel"] <= 1e-3 and res["grad_cosine"] >= 0.999
    # deterministic mode: the same step twice is the same bits (fixed-order BatchNorm statistics, BatchNorm-backward
    # sums, bias and head gradients; split-K wgrad partials summed in order)
    assert res["repeat_bit_identical"], res


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_n_ranks_equal_one_rank_on_the_concatenated_batch(exchange):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (bench.py --gpus N runs this check as its pre-check)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, MSP_SELFCHECK_EXCHANGE=exchange, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "medsegpretrainimagenet_b200.selfcheck"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    assert res["world"] == world and res["ok"], res
    # the exchanges themselves, teacher-forced one layer deep: tight bounds
    ex = res["exchange"]
    assert ex["ok"] and ex["syncbn_y_rms_rel"] <= 1e-4 and ex["syncbn_dx_rms_rel"] <= 1e-4, ex
    assert ex["reduced_dw_rel"] <= 1e-4 and ex["dice_loss_rel"] <= 1e-6 and ex["dice_grad_rel"] <= 1e-5, ex
    # the whole step: as close to the single-device step as that step is to itself under another summation order
    assert res["loss_rel"] <= 1e-3 and res["counters_exchange_bit_exact"]
    assert res["grad_cosine"] >= min(0.999, res["noise_floor_cosine"] - 0.08), res
