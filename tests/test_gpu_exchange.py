"""K-r (otb_grad_exchange): the data-parallel gradient exchange over NVLink peer memory.  Needs two GPUs on the box;
the single-GPU round-end tier skips the multi-rank part and still checks that the entry point refuses bad arguments."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def test_exchange_refuses_bad_arguments_before_touching_the_gpu():
    import ctypes as C
    from onetrainer_b200 import _native as N
    a = N.GradExchangeArgs(n_elems=1024, rank=0, world=1, blocks=1, mode=N.EXCHANGE_P2P)
    with pytest.raises(ValueError, match="rank 0 of 1"):
        N.check(N.lib().otb_grad_exchange(C.byref(a), None), "otb_grad_exchange")
    a = N.GradExchangeArgs(n_elems=1022, rank=1, world=2, blocks=1, mode=N.EXCHANGE_P2P)
    with pytest.raises(ValueError, match="multiple of 4"):
        N.check(N.lib().otb_grad_exchange(C.byref(a), None), "otb_grad_exchange")
    a = N.GradExchangeArgs(n_elems=1024, rank=1, world=2, blocks=1, mode=7)
    with pytest.raises(ValueError, match="mode"):
        N.check(N.lib().otb_grad_exchange(C.byref(a), None), "otb_grad_exchange")


@pytest.mark.gpu
def test_two_rank_exchange_equals_nccl_allreduce_and_norm():
    """Both modes (P2P, NVLS) on 2 ranks: the arena equals ncclAllReduce's bit for bit (a + b has one rounding), every
    replica is identical, and the fused squared-norm partials give ||sum / n||."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29547", str(REPO / "tools" / "exchange_check.py"), str((1 << 22) + 4 * 37)],
                         cwd=REPO, env=env, capture_output=True, text=True, timeout=240)
    assert "exchange_check: ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("equal_to_nccl=True") == 4


def test_shard_exchange_refuses_bad_arguments_before_touching_the_gpu():
    import ctypes as C
    from onetrainer_b200 import _native as N
    a = N.ShardExchangeArgs(begin=0, end=1024, dtype=N.BF16, rank=0, world=1, blocks=1, mode=N.EXCHANGE_P2P)
    with pytest.raises(ValueError, match="rank 0 of 1"):
        N.check(N.lib().otb_shard_reduce(C.byref(a), None), "otb_shard_reduce")
    a = N.ShardExchangeArgs(begin=4, end=1024, dtype=N.BF16, rank=1, world=2, blocks=1, mode=N.EXCHANGE_P2P)
    with pytest.raises(ValueError, match="16-byte vectors"):
        N.check(N.lib().otb_shard_gather(C.byref(a), None), "otb_shard_gather")
    a = N.ShardExchangeArgs(begin=0, end=1024, dtype=99, rank=1, world=2, blocks=1, mode=N.EXCHANGE_NVLS)
    with pytest.raises(ValueError, match="dtype"):
        N.check(N.lib().otb_shard_reduce(C.byref(a), None), "otb_shard_reduce")
    with pytest.raises(ValueError, match="phase"):
        N.check(N.lib().otb_peer_barrier((C.c_void_p * N.MAX_RANKS)(), 0, 2, 9, None), "otb_peer_barrier")


@pytest.mark.gpu
def test_two_rank_sharded_exchange_and_sharded_optimizer_step():
    """K-z on 2 ranks (tools/shard_check.py): otb_shard_reduce equals ncclAllReduce bit for bit on the owner's ranges (bf16
    and fp32, P2P and NVLS), the partial tables give the global norm and are identical on every rank, otb_shard_gather
    makes every parameter replica identical; then the sharded optimizer step (reduce-scatter -> K-c on the owned tensors
    -> all-gather; eager and captured, overlapped and not) reproduces the replicated step: bit for bit without clipping."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29549", str(REPO / "tools" / "shard_check.py")],
                         cwd=REPO, env=env, capture_output=True, text=True, timeout=600)
    assert "shard_check: ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
