"""2+ GPU check of the gradient exchange (run under torchrun): the fused peer-memory all-reduce kernel against the
NCCL all-reduce on the same gradients, eager and inside a CUDA graph, with timings.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeptreeattention_b200 import Hang2020 as H, distributed as D  # noqa: E402
from deeptreeattention_b200.graph import GraphedTrainStep  # noqa: E402
from deeptreeattention_b200.loss import cross_entropy_heads  # noqa: E402


def main():
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    bands, classes, B = 369, 50, 256
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(B, bands, 11, 11, generator=g).to(dev)
    y = torch.randint(0, classes, (B,), generator=g).to(dev)

    def build(peer):
        torch.manual_seed(0)
        m = H.Hang2020(bands, classes).to(dev).train()
        return m, D.GradSync(m, peer=peer)

    def step(m, sync, regime):
        for p in m.parameters():
            p.grad = None
        out = m(x)
        loss = cross_entropy_heads([out] if regime == "R1" else m.head_scores + [out], y)
        loss.backward()
        sync.sync()
        return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    mp, sp = build(True)
    mn, sn = build(False)
    if rank == 0:
        print("peer setup error:", sp.peer_error, "| multicast ptr:", hex(sp.peer.multicast_ptr) if sp.peer else None, flush=True)
    ok = True
    for regime in ("R2+joint", "R1"):
        gp = step(mp, sp, regime)
        gn = step(mn, sn, regime)
        torch.cuda.synchronize()
        worst = 0.0
        bad = []
        for k in gn:
            assert k in gp, k
            scale = float(gn[k].abs().max()) + 1e-12
            d = float((gp[k].double() - gn[k].double()).abs().max()) / scale
            worst = max(worst, d)
            if d > 1e-5:
                bad.append((round(d, 4), k))
        if bad:
            print(f"  rank {rank} tensors that differ:", sorted(bad, reverse=True)[:8], "alpha peer/nccl/buffer slot:",
                  float(gp["alpha"]), float(gn["alpha"]), float(sp.peer.alpha), flush=True)
        # every rank must hold the same averaged gradient bits on the peer path
        flat = mp.fused_spec().flat_grad
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, flat))
        if rank == 0:
            print(f"{regime}: paths {sp.last_path} vs {sn.last_path}; worst relative difference {worst:.2e}; identical across ranks: {same}", flush=True)
        ok = ok and worst < 1e-5 and same
    # timings of the exchange alone
    for name, m, s in (("peer", mp, sp), ("nccl", mn, sn)):
        for _ in range(5):
            s.sync()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            s.sync()
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{name} exchange: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call ({s.last_path})", flush=True)
    # inside a CUDA graph
    mg, sg = build(True)
    gs = GraphedTrainStep(mg, x, y, lambda m, out, yy: cross_entropy_heads(m.head_scores + [out], yy), after_backward=sg.sync)
    gs()
    torch.cuda.synchronize()
    ge = step(mp, sp, "R2+joint")
    torch.cuda.synchronize()
    worst = max(float((mg.get_parameter(k).grad - ge[k]).abs().max()) / (float(ge[k].abs().max()) + 1e-12) for k in ge)
    # the one-call training step (train.fused_train_step, regime R2) followed by the same exchange, eager and captured,
    # against the autograd path on the same crops
    from deeptreeattention_b200.train import GraphedFusedTrainStep, fused_train_step
    ma, sa = build(True)
    for p in ma.parameters():
        p.grad = None
    ma(x)
    cross_entropy_heads(ma.head_scores, y).backward()
    sa.sync()
    mf, sf = build(True)
    fused_train_step(mf, x, y)
    sf.sync()
    mh, sh = build(True)
    gf = GraphedFusedTrainStep(mh, x, y, after_backward=sh.sync)
    gf()
    torch.cuda.synchronize()
    fused_same = all((pa.grad is None) == (pf.grad is None) and (pa.grad is None or (torch.equal(pa.grad, pf.grad) and torch.equal(pa.grad, ph.grad)))
                     for pa, pf, ph in zip(ma.parameters(), mf.parameters(), mh.parameters()))
    if rank == 0:
        print(f"graph replay vs eager (peer path {sg.last_path}): worst relative difference {worst:.2e}", flush=True)
        print(f"fused training step + exchange ({sf.last_path}) == autograd path + exchange, eager and captured, bit for bit: {fused_same}", flush=True)
        print("DIST CHECK", "OK" if ok and worst < 1e-5 and fused_same else "FAILED", flush=True)
    sys.stdout.flush()
    torch.cuda.synchronize(); dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
