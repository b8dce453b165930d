"""CPU tests: the C-ABI library loads, exports every declared symbol, and fails loudly without a GPU;
host-side helpers (synthetic generator, sharding, statistics all-reduce over gloo)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from quadruped_locomotion_b200 import capi, dist as qdist, legmodel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(qlb_built):
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "qlb.h")).read()
    declared = set(re.findall(r"\b(qlb_[a-z_0-9]+)\s*\(", header))
    declared -= {"qlb_status", "qlb_state_status"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qlb_abi_version() == 2


def test_default_params_are_the_reference_gains(qlb_built):
    p = capi.default_params()
    # balance_controller/config/controller_gains.yaml:2-41
    assert list(p.wrench_weights) == [1, 5, 1, 10, 10, 5]
    assert p.ground_force_weight == 1e-4 and p.min_normal_force == 10 and p.friction_default == 0.6
    assert list(p.kp_translation) == [5000, 5000, 10000] and list(p.kd_translation) == [5000, 4000, 5000]
    assert list(p.kp_rotation) == [10000, 10000, 4000] and list(p.kff_rotation) == [0.2, 0.2, 1000]
    assert p.torso_mass == 27.0 and list(p.leg_mass) == [6.0] * 4 and p.gravity == 9.8


def test_struct_layouts_match_header(qlb_built):
    assert C.sizeof(capi.LegModel) == 40 * 8
    assert C.sizeof(capi.Stats) == capi.STATS_NUM * 8
    assert C.sizeof(capi.Params) == (6 + 4 + 18 + 1 + 4 + 12 + 3 + 1 + 1) * 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(qlb_built):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.Solver("quadruped_model")
    lib = capi.load()
    assert lib.qlb_solve_wrench(None, 4, *([None] * 11)) == -4  # QLB_ERR_NOT_INITIALISED
    assert b"CUDA" in lib.qlb_strerror(-2)


def test_models_shipped():
    for name in ("quadruped_model", "simpledog"):
        m = legmodel.load_model(name)
        assert len(m["legs"]) == 4
        assert [leg["link_names"][3] for leg in m["legs"]] == list(legmodel.FOOT_LINKS)
        flat = legmodel.model_to_flat(m)
        assert flat.shape == (4, 40) and np.isfinite(flat).all()
    hdr = open(os.path.join(ROOT, "include", "qlb_models.h")).read()
    assert "QLB_MODEL_QUADRUPED_MODEL" in hdr and "QLB_MODEL_SIMPLEDOG" in hdr and "1.5708" in hdr


def test_synth_is_counter_based():
    a = synth.make_states("C3", 1000)
    b = synth.make_states("C3", 400, start=300)
    for k in a:
        assert np.array_equal(a[k][..., 300:700], b[k]), k
    np.testing.assert_allclose(np.linalg.norm(a["quat"], axis=0), 1.0, atol=1e-14)
    frac4 = (a["mask"] == 0xF).mean()
    assert 0.5 < frac4 < 0.7
    c2 = synth.make_states("C2", 8)
    assert list(c2["mask"]) == [0b1010, 0b0101] * 4
    c1 = synth.make_states("C1")
    assert c1["q"].shape == (12, 1) and c1["mask"][0] == 0xF
    c5 = synth.make_states("C5", 2048)
    assert np.array_equal(c5["q"][:, 0], c5["q"][:, 1023]) and not np.array_equal(c5["q"][:, 0], c5["q"][:, 1024])
    assert c5["mu"].min() >= 0.2 and c5["mu"].max() <= 1.0


def test_shard_ranges_tile_the_batch():
    for B in (1, 7, 1 << 20, 12345):
        for world in (1, 2, 4, 8):
            edges = [qdist.shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank solves its own slice with the oracle and reduces its statistics; no data-path collective
    from oracle import oracle as O
    M = O.model_array(legmodel.load_model("quadruped_model"))
    B = 600
    lo, hi = qdist.shard_range(B, rank, world)
    st = synth.make_states("C3", hi - lo, start=lo)
    r = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"])
    stats = torch.zeros(capi.STATS_NUM, dtype=torch.float64)
    stats[0] = hi - lo
    stats[1] = float((((r["flags"] >> 24) & 7) == 0).sum())
    err = np.sqrt((np.array([1, 5, 1, 10, 10, 5.])[:, None] * (r["netwrench"] - st["wrench"]) ** 2).sum(0))
    stats[7] = float(err.sum())
    stats[29] = float(err.max())
    qdist.allreduce_stats(stats)
    if rank == 0:
        q.put(stats.numpy().copy())
    dist.destroy_process_group()


def test_stats_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process answer over the whole batch
    from oracle import oracle as O
    M = O.model_array(legmodel.load_model("quadruped_model"))
    st = synth.make_states("C3", 600)
    r = O.solve_wrench_batch(M, st["q"], st["quat"], st["wrench"], st["mask"], mu=st["mu"], normals=st["normals"])
    err = np.sqrt((np.array([1, 5, 1, 10, 10, 5.])[:, None] * (r["netwrench"] - st["wrench"]) ** 2).sum(0))
    assert got[0] == 600 and got[1] == 600
    assert abs(got[7] - err.sum()) < 1e-8 * err.sum()
    assert got[29] == err.max()


def test_fused_kernel_has_no_divergence_slow_paths(qlb_built):
    """Build check (no GPU): the fused kernel's SASS stays compact.  When the compiler cannot prove that the warp is
    converged (a box number derived from threadIdx, a trap inside a spin loop, ...) it gives every shuffle an
    out-of-line slow path; the kernel then grows from ~3 700 to ~6 000 instructions and, being instruction-cache
    bound, runs 25 % slower (measured twice in round 2).  tools/sass_stats.py prints the same counts."""
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    txt = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    counts, name = {}, None
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            counts[name] = 0
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            counts[name] += 1
    fused = {k: v for k, v in counts.items() if "qlb_single_kernelIdd" in k}
    assert len(fused) == 4
    for k, v in fused.items():
        assert v < 4800, (k, v)
    tma = [k for k in fused if k.endswith("Lb1EEEvNS_10SolveArgsTIT_EENS_9FusedMapsE")]
    assert tma and all("UTMALDG" in txt for _ in tma)
    # the reciprocals and reciprocal square roots of the solver start from the FP64 seed instructions (no conversion to
    # FP32 and back in the dependent chains), and nothing spills
    body, name = {}, None
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            body[name] = []
        elif name:
            body[name].append(ln)
    for k in fused:
        sass = "\n".join(body[k])
        assert "MUFU.RSQ64H" in sass and "MUFU.RCP64H" in sass, k
        assert not re.search(r"\b(STL|LDL)\b", sass), k
