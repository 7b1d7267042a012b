import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.common import *
g, meta = golden("ref_small_T4")
sd = acoustic_sd()
xin = torch.from_numpy(g["rq_in"])
with torch.no_grad():
    qo, co = O.rq_quantize(xin[None], sd)
print("oracle(on this box) vs golden codes:", (co[0].numpy() == g["rq_codes"]).mean(), "x sum", float(xin.double().sum()), flush=True)
print("codebook0 checksum", float(sd["style_extractor.rqvae.codebooks.0.weight"].double().sum()), float(sd["mel_out.weight"].double().sum()))
print("torch threads", torch.get_num_threads())
# margins on this box
res = xin.clone()
cb = sd["style_extractor.rqvae.codebooks.0.weight"][:-1]
dist = torch.addmm(res.pow(2.).sum(1, keepdim=True) + cb.t().pow(2.).sum(0, keepdim=True), res, cb.t(), alpha=-2.0)
print("depth0 argmin vs golden:", (dist.argmin(1).numpy() == g["rq_codes"][:, 0]).mean())
dist64 = (res.double()**2).sum(1, keepdim=True) + (cb.double()**2).sum(1)[None] - 2 * res.double() @ cb.double().t()
print("depth0 fp64 argmin vs golden:", (dist64.argmin(1).numpy() == g["rq_codes"][:, 0]).mean())
m = acoustic_engine(4)
q, codes = m.rvq(xin.to("cuda:0"), np.array([0, 64], np.int32))
cg = codes.cpu().numpy()
print("gpu vs golden", (cg == g["rq_codes"]).mean(), "gpu vs oracle-here", (cg == co[0].numpy()).mean())
print(cg[:3].tolist(), co[0][:3].tolist(), g["rq_codes"][:3].tolist())
