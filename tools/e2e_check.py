"""End-to-end time of the host entry points (pinned host memory, copies included), FP64 and FP32 interface."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_locomotion_b200 import capi, synth
B = 1 << 20
st = synth.make_states("C3", B)
sol = capi.Solver("quadruped_model", max_batch=B)
for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
    keys = ("q", "quat", "wrench", "mask", "mu")
    h = {k: torch.from_numpy(np.ascontiguousarray(st[k], dtype=(dt if st[k].dtype == np.float64 else st[k].dtype))).pin_memory() for k in keys}
    g = torch.empty((12, B), dtype=tdt).pin_memory(); t = torch.empty((12, B), dtype=tdt).pin_memory()
    n = torch.empty((6, B), dtype=tdt).pin_memory(); f = torch.empty(B, dtype=torch.int32).pin_memory()
    for _ in range(2):
        sol.solve_wrench_host(h["q"], h["quat"], h["wrench"], h["mask"], h["mu"], None, g, t, f, n)
    t0 = time.perf_counter()
    for _ in range(8):
        sol.solve_wrench_host(h["q"], h["quat"], h["wrench"], h["mask"], h["mu"], None, g, t, f, n)
    dtm = (time.perf_counter() - t0) / 8
    print(dt.__name__, "e2e %.3f ms -> %.3e QP/s" % (dtm * 1e3, B / dtm))
