#!/usr/bin/env python
"""End-to-end time of qlb_solve_records_host (pinned host records, copies inside the timed region) - the call bench.py
reports as `e2e`; select a library variant with QLB_LIB."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_locomotion_b200 import capi, synth
B = 1 << 20
st = synth.make_states("C3", B)
sol = capi.Solver("quadruped_model", max_batch=B)
h_rec = torch.from_numpy(capi.wrench_records(st).view(np.uint8).reshape(-1)).pin_memory()
h_res = torch.empty(B * capi.RESULT_RECORD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
for _ in range(3):
    sol.solve_records_host(h_rec, h_res)
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(8):
        sol.solve_records_host(h_rec, h_res)
    best = min(best, (time.perf_counter() - t0) / 8)
print("records e2e %.3f ms -> %.4e QP/s  (%s)" % (best * 1e3, B / best, os.environ.get("QLB_LIB", "in-tree library")))
