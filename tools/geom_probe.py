"""Throughput probe of the geometry kernels (device-resident inputs, CUDA events)."""
import ctypes as C, json, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsbb15_b200 as rg
from tsbb15_b200 import _cabi as cabi
lib = cabi.load_library(); ctx = cabi.context(0)
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tsbb15-3d-reconstruction-project_b200", "data", "dino_data.npz"))
Ps = d["Ps"]
rng = np.random.default_rng(0)
out = {}
for N in (50_000, 1_000_000, 4_000_000):
    Xw = np.column_stack([rng.uniform(-0.045, 0.045, N), rng.uniform(-0.08, 0.03, N), rng.uniform(-0.72, -0.54, N), np.ones(N)])
    def proj(P):
        y = Xw @ P.T
        return y[:, :2] / y[:, 2:]
    a = torch.tensor(proj(Ps[0]) + rng.normal(0, 0.5, (N, 2)), device="cuda")
    b = torch.tensor(proj(Ps[1]) + rng.normal(0, 0.5, (N, 2)), device="cuda")
    C1 = torch.tensor(Ps[0:1].copy(), device="cuda"); C2 = torch.tensor(Ps[1:2].copy(), device="cuda")
    X = torch.empty((N, 3), dtype=torch.float64, device="cuda")
    off = np.array([0, N], dtype=np.int32)
    st = torch.cuda.current_stream().cuda_stream
    for method in (0, 1):
        def call():
            cabi.check(lib.rg_triangulate_dev(C.c_void_p(ctx), C.c_void_p(st), 1, C.c_void_p(C1.data_ptr()), C.c_void_p(C2.data_ptr()),
                       off.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), method, C.c_void_p(X.data_ptr())))
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); 
        for _ in range(5): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"tri_{'optimal' if method == 0 else 'linear'}_N{N}"] = {"ms": ms, "points_per_s": N / ms * 1e3}
print(json.dumps(out, indent=1))
