"""Where does the time of a small batched host call go (BASELINE config 2: 35 Dino pairs x 10 000 hypotheses)?"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsbb15_b200 as rg
from tsbb15_b200 import _cabi as cabi, runtime as rt, sampling, synth
pairs = [np.ascontiguousarray(np.hstack(synth.dino_noisy_pair(i, i + 1))) for i in range(35)]
idl = [sampling.fast(p.shape[0], 10000, 8, seed=i) for i, p in enumerate(pairs)]
rt.f_ransac_batched(pairs, idl, thr=1.5)
def t(fn, n=20):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print("python wrapper total ms", t(lambda: rt.f_ransac_batched(pairs, idl, thr=1.5)))
print("  without mask       ms", t(lambda: rt.f_ransac_batched(pairs, idl, thr=1.5, want_mask=False)))
# raw C call with preconcatenated arrays
lib = cabi.load_library(); ctx = cabi.context(0); vp = C.c_void_p
pts = np.ascontiguousarray(np.concatenate(pairs)); idx = np.ascontiguousarray(np.concatenate(idl))
po = np.zeros(36, np.int32); ho = np.zeros(36, np.int32)
po[1:] = np.cumsum([p.shape[0] for p in pairs]); ho[1:] = np.cumsum([i.shape[0] for i in idl])
bi = np.zeros(35, np.int32); bc = np.zeros(35, np.int32); bF = np.zeros((35, 9)); mk = np.zeros(len(pts), np.uint8)
pi = C.POINTER(C.c_int32)
def raw(): cabi.check(lib.rg_f_ransac_host(vp(ctx), None, 35, vp(pts.ctypes.data), po.ctypes.data_as(pi), vp(idx.ctypes.data), ho.ctypes.data_as(pi), 1.5, 0, 0, 0, 0, vp(bi.ctypes.data), vp(bc.ctypes.data), vp(bF.ctypes.data), vp(mk.ctypes.data), None, None, None))
print("raw C host call     ms", t(raw))
tp = torch.from_numpy(pts).pin_memory(); ti = torch.from_numpy(idx).pin_memory()
tbi = torch.zeros(35, dtype=torch.int32).pin_memory(); tbc = torch.zeros(35, dtype=torch.int32).pin_memory(); tbF = torch.zeros(35, 9, dtype=torch.float64).pin_memory(); tmk = torch.zeros(len(pts), dtype=torch.uint8).pin_memory()
def rawp(): cabi.check(lib.rg_f_ransac_host(vp(ctx), None, 35, vp(tp.data_ptr()), po.ctypes.data_as(pi), vp(ti.data_ptr()), ho.ctypes.data_as(pi), 1.5, 0, 0, 0, 0, vp(tbi.data_ptr()), vp(tbc.data_ptr()), vp(tbF.data_ptr()), vp(tmk.data_ptr()), None, None, None))
print("raw C, pinned       ms", t(rawp))
dp = tp.cuda(); di = ti.cuda(); dbi = torch.zeros(35, dtype=torch.int32, device="cuda"); dbc = torch.zeros_like(dbi); dbF = torch.zeros(35, 9, dtype=torch.float64, device="cuda"); dmk = torch.zeros(len(pts), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def dev():
    cabi.check(lib.rg_f_ransac_dev(vp(ctx), vp(st), 35, vp(dp.data_ptr()), po.ctypes.data_as(pi), vp(di.data_ptr()), ho.ctypes.data_as(pi), 1.5, 0, 0, 0, 0, vp(dbi.data_ptr()), vp(dbc.data_ptr()), vp(dbF.data_ptr()), vp(dmk.data_ptr())))
def devs(): dev(); torch.cuda.synchronize()
print("dev call + sync     ms", t(devs))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): dev()
e1.record(); torch.cuda.synchronize()
print("dev call GPU time   ms", e0.elapsed_time(e1) / 20)
rt.set_option(1, 1); dev(); dev(); print(rt.profile(stream=st)); rt.set_option(1, 0)
print("bytes in", pts.nbytes + idx.nbytes, "evals", sum(p.shape[0] for p in pairs) * 10000)
