#!/usr/bin/env python
"""Build the CPU oracle (TEST INFRASTRUCTURE ONLY).

  oracle/liboracle.so            our C restatement (always)
  oracle/_ref/libquadprog_ref.so the reference's own QuadProg++ compiled IN PLACE from
                                 /root/reference (only where that tree exists; the built .so
                                 travels to the GPU box, the sources are never copied)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QLB_REFERENCE", "/root/reference")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(verbose=False, force=False):
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("qlb_oracle.c", "qlb_oracle_ipm.c", "qlb_oracle_vmc.c")]
    deps = srcs + [os.path.join(HERE, "qlb_oracle.h")]
    if force or _stale(out, deps):
        cmd = ["gcc", "-O2", "-std=c11", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas",
               "-ffp-contract=off", "-o", out] + srcs + ["-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    ref_out = os.path.join(HERE, "_ref", "libquadprog_ref.so")
    ref_src = [os.path.join(REF, "qp_solver", "src", f) for f in ("QuadProg++.cc", "Array.cc")]
    if all(os.path.exists(s) for s in ref_src):
        shim = os.path.join(HERE, "ref_shim.cc")
        if force or _stale(ref_out, ref_src + [shim]):
            os.makedirs(os.path.dirname(ref_out), exist_ok=True)
            cmd = ["g++", "-std=c++14", "-O2", "-w", "-shared", "-fPIC",
                   "-I" + os.path.join(REF, "qp_solver", "include"), "-o", ref_out, shim] + ref_src
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    return out, (ref_out if os.path.exists(ref_out) else None)


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
