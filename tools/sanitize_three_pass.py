#!/usr/bin/env python
"""Repeated three-pass solves with F_min = 0 on the C5 test batch (run under compute-sanitizer)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadruped_locomotion_b200 import capi, synth
st = synth.make_states("C5", 8192, start=999)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ref = None
for t in range(n):
    s = capi.Solver("quadruped_model")
    s.set_pipeline(sys.argv[1] if len(sys.argv) > 1 else "three_pass")
    p = s.get_params(); p.min_normal_force = 0.0; s.set_params(p)
    out = s.solve_wrench_numpy(st)
    if ref is None: ref = out
    d = np.abs(out["grf"] - ref["grf"]).max()
    print("trial", t, "ok", int((((out["flags"] >> 24) & 7) == 0).sum()), "max diff vs first %.2e" % d, flush=True)
