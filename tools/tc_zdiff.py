"""Forward only: compare the saved conv outputs z1/z2 of conv_impl 1 vs 0.  Usage: python tools/tc_zdiff.py kind bands classes batch dist"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeptreeattention_b200 import Hang2020 as H, _capi
from oracle import hang2020_oracle as orc
CLS = {"hang2020": H.Hang2020, "spectral": H.spectral_network, "spatial": H.spatial_network}
kind, bands, classes, batch, dist = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
seed = 1000 + batch
table = orc.init_params(kind, bands, classes, seed, perturb_bn=True)
x, y = orc.make_inputs(batch, bands, classes, seed, dist)
nb = 2 if kind == "hang2020" else 1
zs = []
for impl in (0, 1):
    _capi.set_option(0, "conv_impl", impl)
    m = CLS[kind](bands, classes); m.load_state_dict(table); m = m.cuda().train()
    out = m(x.cuda())
    o = out if torch.is_tensor(out) else out[0]
    saved = o.grad_fn.saved_tensors[1]
    f = saved.view(torch.float32)
    n1 = batch * nb * 32 * 121
    z1 = f[:n1].view(batch, nb * 32, 121).clone()
    off = (n1 * 4 + 255) // 256 * 256 // 4
    n2 = batch * nb * 64 * 121
    z2 = f[off:off + n2].view(batch, nb * 64, 121).clone()
    zs.append((z1, z2))
    torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x, table["conv1.conv_layer.weight" if nb == 1 else "spectral_network.conv1.conv_layer.weight"],
                                table["conv1.conv_layer.bias" if nb == 1 else "spectral_network.conv1.conv_layer.bias"], padding=1).reshape(batch, 32, 121)
for name, a, b in (("z1", zs[0][0], zs[1][0]), ("z2", zs[0][1], zs[1][1])):
    d = (a - b).abs()
    print(name, "max|tc-simt|", float(d.max()), "scale", float(a.abs().max()), "at", torch.nonzero(d == d.max())[0].tolist())
    print("   per-position max:", [f"{v:.1e}" for v in d.amax(dim=(0, 1)).tolist()][:24])
print("z1 simt vs torch conv:", float((zs[0][0][:, :32].cpu() - ref).abs().max()), " tc vs torch conv:", float((zs[1][0][:, :32].cpu() - ref).abs().max()))
