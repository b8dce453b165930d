#!/usr/bin/env python
"""Quick GPU sanity/diagnostics run: parity vs the CPU oracle + kernel timing. Not a test, not a bench."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402
from quadruped_locomotion_b200 import capi, legmodel, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--model", default="quadruped_model")
    ap.add_argument("--time-batch", type=int, default=1 << 20)
    ap.add_argument("--f32", action="store_true", help="run the _f32 twins")
    args = ap.parse_args()
    st = synth.make_states(args.config, args.batch)
    M = O.model_array(legmodel.load_model(args.model))
    t = time.time()
    ref = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"],
                               solver=O.SOLVER_GI, want_margin=True)
    print("oracle GI: %.2fs" % (time.time() - t))
    sol = capi.Solver(args.model)
    if os.environ.get("QLB_F32_CORE") == "float":
        sol.set_f32_core(False)
    npdt = np.float32 if args.f32 else np.float64
    tdt = torch.float32 if args.f32 else torch.float64
    out = sol.solve_wrench_numpy(st, dtype=npdt)
    sc = np.maximum(1.0, np.abs(ref["grf"]).max(0))
    e = np.abs(out["grf"] - ref["grf"]).max(0) / sc
    sct = np.maximum(1.0, np.abs(ref["tau"]).max(0))
    et = np.abs(out["tau"] - ref["tau"]).max(0) / sct
    en = np.abs(out["netwrench"] - ref["netwrench"]).max(0) / np.maximum(1.0, np.abs(ref["netwrench"]).max(0))
    print("grf rel err: max %.3e median %.3e   tau: max %.3e   net: max %.3e" % (np.nanmax(e), np.nanmedian(e), np.nanmax(et), np.nanmax(en)))
    print("nan outputs:", int(np.isnan(out["grf"]).any(0).sum()))
    fm = (out["flags"] & 0xFFFFFF) != (ref["flags"] & 0xFFFFFF)
    print("flag mismatches:", int(fm.sum()), "of", args.batch)
    stt = (out["flags"] >> 24) & 7
    it = out["flags"] >> 27
    print("status hist", np.bincount(stt, minlength=6), "iters hist", np.bincount(it))
    if args.f32:
        for thr in (1e-2, 1e-3, 1e-4):
            ok = ref["margin"] > thr
            print("  margin > %g: %d states, flag mismatches %d, grf err max %.3e, tau err max %.3e" % (thr, int(ok.sum()), int((fm & ok).sum()), np.nanmax(e[ok]), np.nanmax(et[ok])))
        print("  grf err percentiles 50/90/99/99.9/100:", np.percentile(e, [50, 90, 99, 99.9, 100]))
    bad = np.nonzero((e > (1e-3 if args.f32 else 1e-6)) | (fm & (ref["margin"] > (1e-3 if args.f32 else 0))))[0][:10]
    for i in bad:
        print("  bad", i, "err %.3e" % e[i], "flags gpu %08x ref %08x" % (out["flags"][i], ref["flags"][i]), "margin %.2e" % ref["margin"][i])
    # timing on device-resident inputs
    Bt = args.time_batch
    st = synth.make_states(args.config, Bt)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in st.items()}
    if args.f32:
        d = {k: (v.to(torch.float32) if v.dtype == torch.float64 else v) for k, v in d.items()}
    grf = torch.empty((12, Bt), dtype=tdt, device=dev)
    tau = torch.empty_like(grf)
    flags = torch.empty(Bt, dtype=torch.int32, device=dev)
    net = torch.empty((6, Bt), dtype=tdt, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        sol.solve_wrench(d["q"], d["quat"], d["wrench"], d["mask"], d["mu"], d["normals"], grf, tau, flags, net, stream=stream)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    ev0.record()
    for _ in range(reps):
        sol.solve_wrench(d["q"], d["quat"], d["wrench"], d["mask"], d["mu"], d["normals"], grf, tau, flags, net, stream=stream)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print("device-resident: %.3f ms per %d QPs -> %.3e QP/s" % (ms, Bt, Bt / ms * 1e3))
    s = sol.batch_stats(flags, d["wrench"].double(), net.double(), stream=stream)
    print("stats: count %d status %s mean it %.3f mean err %.4f max err %.4f" % (s[0], s[1:6], s[6] / s[0], s[7] / s[0], s[29]))


if __name__ == "__main__":
    main()
