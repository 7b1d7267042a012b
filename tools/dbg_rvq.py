import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.common import *
g, meta = golden("ref_small_T4")
m = acoustic_engine(4)
sd = acoustic_sd()
x = torch.from_numpy(g["rq_in"])
for trial, xx, offs in [("b1", x, [0, 64]), ("b1again", x, [0, 64]), ("x3", torch.cat([x, x, x]), [0, 64, 128, 192]),
                        ("rand64", torch.randn(64, 256) * 0.8, [0, 64])]:
    q, codes = m.rvq(xx.to("cuda:0").contiguous(), np.array(offs, np.int32))
    with torch.no_grad():
        qo, co = O.rq_quantize(xx[None], sd)
    eq = (codes.cpu().numpy() == co[0].numpy())
    print(trial, "codes match frac", eq.mean(), "per depth", eq.mean(0), "q err", float((q.cpu() - qo[0]).abs().max()))
    if trial == "b1":
        print(codes[:3].cpu().tolist(), co[0][:3].tolist())
