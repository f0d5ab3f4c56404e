"""Run one seeded case repeatedly under a conv_impl and print run-to-run gradient differences (should be 0:
every kernel is deterministic) and the distance to the oracle.  Usage: python tools/tc_repeat.py kind bands classes batch dist regime impl"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.tc_compare import run  # noqa: E402
from oracle import hang2020_oracle as orc  # noqa: E402

kind, bands, classes, batch, dist, regime, impl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6], int(sys.argv[7])
seed = 1000 + batch
table = orc.init_params(kind, bands, classes, seed, perturb_bn=True)
x, y = orc.make_inputs(batch, bands, classes, seed, dist)
_, _, rheads, rgrads, _ = orc.step(kind, table, x, y, regime=regime, training=True)
prev = None
for it in range(3):
    junk = torch.full((64 * 1024 * 1024,), float(it + 1) * 1e3, device="cuda")   # dirty the allocator's free blocks
    del junk
    h, g = run(kind, bands, classes, table, x, y, regime, impl)
    worst = max(((float((g[k] - rgrads[k]).abs().max()) / (float(rgrads[k].abs().max()) + 1e-12)), k) for k in g)
    line = f"impl {impl} run {it}: worst rel grad err vs oracle {worst[0]:.3e} ({worst[1]})"
    if prev is not None:
        d = max((float((g[k] - prev[k]).abs().max()), k) for k in g)
        line += f"   run-to-run max diff {d[0]:.3e} ({d[1]})"
    print(line)
    prev = g
