import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.common import *
g, meta = golden("ref_small_T4")
m = acoustic_engine(4)
DEV = "cuda:0"
Fr = g["dn_spec"].shape[1]
offs = np.array([0, Fr], np.int32)
cond = torch.from_numpy(g["dn_cond"].T.copy()).to(DEV)
x = torch.from_numpy(g["rq_in"]).to(DEV)
def rv(tag):
    q, codes = m.rvq(x, np.array([0, x.shape[0]], np.int32))
    eq = (codes.cpu().numpy().astype(np.int64) == g["rq_codes"]).mean()
    print(tag, "codes match frac", eq, flush=True)
rv("fresh")
e = m.denoiser_eval(0, torch.from_numpy(g["dn_spec"].T.copy()).to(DEV), None, 3, cond, offs); rv("after mel eval")
f0 = torch.from_numpy(g["dd_f0"]).to(DEV); uv = torch.from_numpy(g["dd_uv"].astype(np.int32)).to(DEV)
e2 = m.denoiser_eval(1, f0, uv, 1, cond, offs); rv("after f0 eval 1")
e3 = m.denoiser_eval(2, f0, uv, 0, cond, offs); rv("after f0 eval 2")
rv("again")
m.set_tensor_cores(False)
e3 = m.denoiser_eval(2, f0, uv, 0, cond, offs); rv("after simt f0 eval 2")
e = m.denoiser_eval(0, torch.from_numpy(g["dn_spec"].T.copy()).to(DEV), None, 3, cond, offs); rv("after simt mel eval")
