"""CPU / gloo, world_size 2: the data-parallel plumbing of the path -- sharding of the crops and the
single gradient exchange with DDP semantics (per-rank BatchNorm statistics, gradients averaged over ranks).  The
module under the exchange is the CPU oracle (the CUDA modules cannot run here); the GPU fast path of GradSync (one
all-reduce of the flat gradient buffer) is exercised by bench.py --gpus N on the box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import multiprocessing as mp

from deeptreeattention_b200 import distributed as D
from oracle import hang2020_oracle as orc

KIND, BANDS, CLASSES, GLOBAL_B = "hang2020", 12, 5, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_grads(rank, world):
    torch.set_num_threads(1)
    m = orc.OracleModule(KIND, BANDS, CLASSES, seed=3).train()
    x, y = orc.make_inputs(GLOBAL_B, BANDS, CLASSES, seed=3)
    lo, hi = D.shard_range(GLOBAL_B, rank, world)
    out = m(x[lo:hi])
    loss = sum(torch.nn.functional.cross_entropy(h, y[lo:hi]) for h in m.heads) + torch.nn.functional.cross_entropy(out, y[lo:hi])
    loss.backward()
    return m


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    m = _local_grads(rank, world)
    sync = D.GradSync(m)
    sync.sync()
    D.broadcast_buffers(m, src=0)
    q.put((rank, sync.last_path, {k: p.grad.numpy().copy() for k, p in m.named_parameters()},
           {k: b.numpy().copy() for k, b in m.named_buffers()}))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    for total, world in ((10, 2), (8192, 8), (7, 3), (2, 4)):
        spans = [D.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


@pytest.mark.timeout(180)
def test_grad_sync_world2_gloo_matches_manual_average():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=150) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    # expected: each rank's own gradient (per-rank BN statistics), averaged
    expect = None
    for r in range(world):
        g = {k: p.grad for k, p in _local_grads(r, world).named_parameters()}
        expect = g if expect is None else {k: expect[k] + g[k] for k in g}
    expect = {k: v / world for k, v in expect.items()}
    for rank, path, grads, bufs in results:
        assert path == "generic"
        for k, v in expect.items():
            got = torch.from_numpy(grads[k])
            assert got.dtype == v.dtype
            torch.testing.assert_close(got, v, rtol=1e-6, atol=1e-7)
    # DDP broadcast_buffers: rank 0's running statistics everywhere
    for k, b in results[0][3].items():
        assert (b == results[1][3][k]).all(), k


def test_single_process_is_a_noop():
    m = orc.OracleModule("vanilla", 3, 2, seed=0)
    s = D.GradSync(m)
    s.sync()
    assert s.last_path == "single"
