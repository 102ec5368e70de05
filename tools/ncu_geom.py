"""Small driver for ncu: a few launches of the geometry kernels on device-resident data."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import _cabi as cabi
lib = cabi.load_library(); ctx = cabi.context(0)
vp = C.c_void_p
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tsbb15-3d-reconstruction-project_b200", "data", "dino_data.npz"))
Ps = d["Ps"]
rng = np.random.default_rng(0)
N = 1_000_000
Xw = np.column_stack([rng.uniform(-0.045, 0.045, N), rng.uniform(-0.08, 0.03, N), rng.uniform(-0.72, -0.54, N), np.ones(N)])
proj = lambda P: (Xw @ P.T)[:, :2] / (Xw @ P.T)[:, 2:]
a = torch.tensor(proj(Ps[0]) + rng.normal(0, 0.5, (N, 2)), device="cuda")
b = torch.tensor(proj(Ps[1]) + rng.normal(0, 0.5, (N, 2)), device="cuda")
C1 = torch.tensor(Ps[0:1].copy(), device="cuda"); C2 = torch.tensor(Ps[1:2].copy(), device="cuda")
X = torch.empty((N, 3), dtype=torch.float64, device="cuda")
off = np.array([0, N], dtype=np.int32)
st = torch.cuda.current_stream().cuda_stream
for method in (0, 0, 0, 1, 1):
    cabi.check(lib.rg_triangulate_dev(vp(ctx), vp(st), 1, vp(C1.data_ptr()), vp(C2.data_ptr()), off.ctypes.data_as(C.POINTER(C.c_int32)),
                                      vp(a.data_ptr()), vp(b.data_ptr()), method, vp(X.data_ptr())))
M = 20000
obs = torch.tensor(np.column_stack([rng.uniform(-0.1, 0.1, (M, 2)), np.ones(M)]), device="cuda")
idx = torch.empty(M, dtype=torch.int32, device="cuda")
for _ in range(2):
    cabi.check(lib.rg_match_first_within_dev(vp(ctx), vp(st), 3, M, vp(obs.data_ptr()), M, vp(obs.data_ptr()), 1e-4, vp(idx.data_ptr())))
torch.cuda.synchronize()
print("ok", float(X.abs().max()), int((idx >= 0).sum()))
