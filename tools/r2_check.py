"""Round-2 bring-up check on a B200: device Philox vs host twin, seeded calls vs host-drawn indices, flag-list overflow
fallback, multi-pass batches, PnP solvers, and first timings.  Prints one JSON line per section."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import tsbb15_b200 as rg  # noqa: E402
from tsbb15_b200 import device as dv, philox, runtime as rt, sampling, synth  # noqa: E402


def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {}
    # 1. philox: device == host
    pts_d, cams_d = dv.synth_two_view(3, 5000, first_pair=11)
    ok = True
    for p in range(3):
        h, cp = philox.synth_two_view(5000, 11 + p, synth.dino()["Ps"], synth.DINO_BBOX)
        ok &= bool(np.array_equal(pts_d[p].cpu().numpy(), h)) and tuple(cams_d[p].cpu().numpy()) == cp
    idx_d = dv.sample_indices([5000, 300, 8], [700, 50, 20], 8, seed=99, first_pair=4, hyp_first=3).cpu().numpy()
    idx_h = np.concatenate([philox.sample_indices(5000, 700, 8, 99, 4, 3), philox.sample_indices(300, 50, 8, 99, 5, 3),
                            philox.sample_indices(8, 20, 8, 99, 6, 3)])
    out["philox"] = {"synth_equal": ok, "idx_equal": bool(np.array_equal(idx_d, idx_h))}
    print(json.dumps(out), flush=True)

    # 2. seeded host call == host-drawn indices call; multi-pass == single pass; overflow fallback == normal
    pairs = [pts_d[p].cpu().numpy() for p in range(3)]
    r_seed = rt.f_ransac_batched(pairs, None, n_hyp=600, sample_seed=7, first_pair=2, want_counts=True)
    idl = philox.sample_indices_batch([5000] * 3, 600, 8, 7, 2)
    r_host = rt.f_ransac_batched(pairs, idl, want_counts=True)
    same = all(np.array_equal(a, b) for a, b in zip(r_seed["counts"], r_host["counts"])) and \
        np.array_equal(r_seed["best_idx"], r_host["best_idx"]) and np.array_equal(r_seed["F"], r_host["F"])
    rt.set_option(6, 5000 * 600)                      # one pair per pass
    r_mp = rt.f_ransac_batched(pairs, idl)
    st_mp = rt.last_stats()
    rt.set_option(6, 0)
    rt.set_option(8, 64)                              # tiny flag list: FP64 recount fallback
    r_ov = rt.f_ransac_batched(pairs, idl, want_counts=True)
    st_ov = rt.last_stats()
    rt.set_option(8, 0)
    r_64 = rt.f_ransac_batched(pairs, idl, want_counts=True, score_path=rg.SCORE_FP64)
    out["f_path"] = {
        "seeded_equals_host_drawn": bool(same),
        "multipass_equal": bool(np.array_equal(r_mp["best_idx"], r_host["best_idx"]) and np.array_equal(r_mp["F"], r_host["F"])
                                and all(np.array_equal(a, b) for a, b in zip(r_mp["mask"], r_host["mask"]))),
        "multipass_passes": st_mp["passes"],
        "overflow_equal": bool(all(np.array_equal(a, b) for a, b in zip(r_ov["counts"], r_host["counts"]))),
        "overflow_recounted": st_ov["overflow"],
        "fp32_guarded_equals_fp64": bool(all(np.array_equal(a, b) for a, b in zip(r_64["counts"], r_host["counts"]))),
        "best_count": r_host["best_count"].tolist()}
    print(json.dumps(out["f_path"]), flush=True)

    # 3. PnP: new solver vs group Jacobi
    X, y, _ = synth.pnp_scene(20000, seed=4)
    pidx = sampling.fast(20000, 4096, 6, seed=2)
    thr2 = (1.5 / 3217.0) ** 2
    a = rt.pnp_ransac(X, y, pidx, thr2, want_counts=True, want_poses=True, want_flags=True)
    rt.set_option(7, 1)
    b = rt.pnp_ransac(X, y, pidx, thr2, want_counts=True, want_poses=True, want_flags=True)
    rt.set_option(7, 0)
    good = (a["flags"] == 0) & (b["flags"] == 0)
    dpose = np.abs(a["poses"][good] - b["poses"][good]).max(axis=1)
    out["pnp"] = {"counts_equal_unflagged": int((a["counts"][good] != b["counts"][good]).sum()), "n_unflagged": int(good.sum()),
                  "flag_disagree": int((a["flags"] != b["flags"]).sum()),
                  "pose_diff_median": float(np.median(dpose)), "pose_diff_max": float(dpose.max()),
                  "best": [a["best_idx"], b["best_idx"], a["best_count"], b["best_count"]]}
    print(json.dumps(out["pnp"]), flush=True)
    dev = torch.device("cuda", 0)
    N4, H4 = 1000000, 8192
    X4, y4, _ = synth.pnp_scene(N4, seed=4)
    pi4 = sampling.fast(N4, H4, 6, seed=2)
    dX, dy, dI = (torch.from_numpy(v).to(dev) for v in (X4, y4, pi4))
    po = dv.PnpOutputs(1, N4, want_mask=True)
    vo, ho = np.array([0, N4], np.int32), np.array([0, H4], np.int32)
    t = {}
    for solver in (0, 1):
        rt.set_option(7, solver)
        rt.set_option(1, 1)
        ms = ev_time(lambda: dv.pnp_ransac(dX, dy, vo, dI, ho, po, thr2), 5)
        pr = rt.profile(stream=torch.cuda.current_stream().cuda_stream)
        rt.set_option(1, 0)
        t["solver%d" % solver] = {"ms": ms, "solve_ms": pr["solve_ms"] / max(pr["calls"], 1), "score_ms": pr["score_ms"] / max(pr["calls"], 1),
                                  "fixup_ms": pr["fixup_ms"] / max(pr["calls"], 1), "count": int(po.best_count.item())}
    rt.set_option(7, 0)
    print(json.dumps({"pnp_config4": t}), flush=True)

    # 4. timing: 16 and 64 pairs of the config-5 shape, device resident, seeded
    for P in (16, 64):
        d_pts, _ = dv.synth_two_view(P, 50000, first_pair=0)
        o = dv.FOutputs(P, P * 50000, want_mask=True)
        po_, ho_ = dv.offsets(np.full(P, 50000)), dv.offsets(np.full(P, 8192))
        rt.set_option(1, 1)
        ms = ev_time(lambda: dv.f_ransac(d_pts, po_, None, ho_, o, seed=5), 5)
        pr = rt.profile(stream=torch.cuda.current_stream().cuda_stream)
        rt.set_option(1, 0)
        st = rt.last_stats(stream=torch.cuda.current_stream().cuda_stream)
        c = max(pr["calls"], 1)
        print(json.dumps({"config5_pairs": P, "ms": ms, "evals_per_s": P * 50000 * 8192 / (ms * 1e-3),
                          "phases": {k: pr[k] / c for k in pr if k != "calls"}, "stats": st,
                          "min_count": int(o.best_count.min().item())}), flush=True)


if __name__ == "__main__":
    main()
