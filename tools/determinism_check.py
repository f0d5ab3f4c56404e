import sys, torch
sys.path.insert(0, "/root/repo")
from deeptreeattention_b200 import Hang2020 as H
from oracle import hang2020_oracle as orc
for (bands, classes, B) in ((30, 6, 16), (369, 50, 64), (369, 50, 1024)):
    table = orc.init_params("hang2020", bands, classes, 9)
    x, y = orc.make_inputs(B, bands, classes, 9)
    xd, yd = x.cuda(), y.cuda()
    ref = None
    for rep in range(6):
        m = H.Hang2020(bands, classes); m.load_state_dict(table); m = m.cuda().train()
        loss = torch.nn.functional.cross_entropy(m(xd), yd); loss.backward(); torch.cuda.synchronize()
        g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        if ref is None: ref = g
        else:
            bad = [k for k in g if not torch.equal(g[k], ref[k])]
            print((bands, classes, B), "rep", rep, "differing tensors:", bad[:6])
