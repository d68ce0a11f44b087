"""Data-parallel host logic on CPU: world_size 2 over gloo (the N>1 path of SURVEY.md §8e).
The path shards by batch with ONE exchange step; these tests cover the exchange (in-place arena
allreduce and the gather/scatter bucket), the 1/n prescale, and the global-sample-index RNG keying
that makes an N-rank run draw what a 1-rank run at the same global batch draws."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from onetrainer_b200 import dp
    from oracle import otb_oracle as O
    rk, lr, ws = dp.init_from_env("gloo")
    assert (rk, ws) == (rank, world) and dp.rank() == rank and dp.world_size() == world
    b = 3
    assert dp.global_sample_base(b) == rank * b
    # (1) arena mode: gradients are views of one flat buffer -> allreduce in place, no copies
    arena = torch.zeros(10)
    g1, g2 = arena[0:4].view(2, 2), arena[4:10]
    g1.fill_(float(rank + 1)); g2.fill_(10.0 * (rank + 1))
    pre = dp.GradientBucket(arena).allreduce([g1, g2])
    assert pre == 1.0 / world
    assert torch.equal(g1, torch.full((2, 2), 3.0)) and torch.equal(g2, torch.full((6,), 30.0))
    # (2) bucket mode: unrelated tensors are gathered, summed, scattered back
    t1, t2 = torch.full((5,), float(rank)), torch.full((2, 3), 2.0 + rank)
    pre = dp.GradientBucket(None).allreduce([t1, t2])
    assert pre == 0.5 and torch.equal(t1, torch.full((5,), 1.0)) and torch.equal(t2, torch.full((2, 3), 5.0))
    # (2b) full fine-tuning: persistent flat gradients (dp.FlatGradArena) -> autograd accumulates into the views in
    # place, the exchange runs in place on the arena, and one memset replaces zero_grad(set_to_none=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    fa = dp.FlatGradArena(list(lin.parameters()))
    ptrs = [p.grad.data_ptr() for p in lin.parameters()]
    for _ in range(2):                                   # two micro-batches accumulate
        lin(torch.full((4, 6), float(rank + 1))).sum().backward()
    assert [p.grad.data_ptr() for p in lin.parameters()] == ptrs
    assert all(dp._is_view_of(p.grad, fa.flat) for p in lin.parameters())
    pre = dp.GradientBucket(fa.flat).allreduce(fa.grads())
    assert pre == 0.5
    assert torch.equal(lin.bias.grad, torch.full((3,), 2 * 4.0 * world))                    # sum over ranks of 2 x 4 ones
    assert torch.equal(lin.weight.grad, torch.full((3, 6), 2 * 4.0 * (1.0 + 2.0)))
    fa.zero()
    assert all(not p.grad.any() for p in lin.parameters()) and [p.grad.data_ptr() for p in lin.parameters()] == ptrs
    # (2c) the same exchange issued bucket by bucket from post-accumulate-grad hooks while backward is still running
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(8, 8) for _ in range(5)])
    fa2 = dp.FlatGradArena(list(net.parameters()))
    red = dp.OverlappedArenaReducer(fa2, n_buckets=3)
    assert len(red._slices) == 3 and sum(sl.numel() for sl in red._slices) == fa2.flat.numel()
    red.arm(last_micro_step=False)                       # first micro-batch of two: accumulate only
    net(torch.full((2, 8), 0.5 + rank)).sum().backward()
    assert red.finish() == 1.0 and red._issued == 0
    local = fa2.flat.clone()
    red.arm()
    net(torch.full((2, 8), 0.5 + rank)).sum().backward()
    assert red._issued >= 2                              # buckets went out during backward, before finish()
    assert red.finish() == 0.5 and red._issued == 3
    gathered = [torch.zeros_like(local) for _ in range(world)]
    torch.distributed.all_gather(gathered, 2 * local)    # what each rank held locally after both micro-batches
    assert torch.allclose(fa2.flat, sum(gathered), rtol=1e-6, atol=1e-6)
    # (2d) draws that must be identical on every replica (stochastic rounding by the stand-alone helpers): rank 0's seed and
    # a call counter, whatever this rank seeded torch with
    torch.manual_seed(100 + rank)
    keys = [dp.replicated_key() for _ in range(3)]
    assert [k[1] for k in keys] == [4, 8, 12] and len({k[0] for k in keys}) == 1
    got = [None] * world
    torch.distributed.all_gather_object(got, keys)
    assert all(g == got[0] for g in got) and keys[0][0] == 100                          # rank 0's seed everywhere
    # (3) loss averaging for logging
    assert dp.average_scalar(torch.tensor(float(rank))).item() == 0.5
    # (4) DP-invariant draws: this rank's samples of the global batch, through the oracle's stream definition
    ts = [O.timestep(O.TS_UNIFORM, 42, 0, dp.global_sample_base(b) + i) for i in range(b)]
    noise = O.create_noise(b, 4, 16, seed=42, offset=0, sample_base=dp.global_sample_base(b))
    np.savez(Path(out_dir) / f"rank{rank}.npz", ts=np.array(ts), noise=noise)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_exchange_and_rng_invariance(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b = 3
    single_ts = [oracle.timestep(oracle.TS_UNIFORM, 42, 0, i) for i in range(world * b)]
    single_noise = oracle.create_noise(world * b, 4, 16, seed=42, offset=0, sample_base=0)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert list(d["ts"]) == single_ts[r * b:(r + 1) * b]
        assert np.array_equal(d["noise"], single_noise[r * b:(r + 1) * b])


def test_single_process_defaults():
    from onetrainer_b200 import dp
    assert dp.world_size() == 1 and dp.rank() == 0 and dp.global_sample_base(8) == 0
    g = torch.ones(4)
    assert dp.GradientBucket(None).allreduce([g]) == 1.0 and torch.equal(g, torch.ones(4))


@pytest.mark.parametrize("n", [2])
def test_bench_reference_arm_other_ranks_exit_without_work(n):
    """Contract: under torchrun only rank 0 runs the CPU reference arm; the others exit 0 silently."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE=str(n), LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", str(n), "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
