"""The device-side state generator (qlb_generate_states) against the host generator (synth.make_states):
bit-identical, for every BASELINE config, any slice of the stream, FP64 and the FP32 twin."""
import numpy as np
import pytest

from quadruped_locomotion_b200 import synth


def test_exact_elementary_functions_are_accurate():
    """CPU: the IEEE-exact sin / cos / log of the generator agree with libm to rounding level."""
    rng = np.random.default_rng(7)
    x = rng.uniform(-30.0, 30.0, 100000)
    s, c = synth.sincos_exact(x)
    assert np.abs(s - np.sin(x)).max() <= 4e-16 and np.abs(c - np.cos(x)).max() <= 4e-16
    u = np.concatenate([rng.uniform(0.0, 1.0, 100000), [1.0, 2.0 ** -53, 0.5, 0.70710678118654752440]])
    l = synth.log_exact(u)
    assert l[100000] == 0.0
    assert np.abs(l - np.log(u)).max() <= 1e-15 * np.maximum(1.0, np.abs(np.log(u))).max()


def test_generator_is_a_pure_function_of_the_index():
    """CPU: a slice generated on its own equals the same slice of a larger batch (rank sharding relies on it)."""
    for cfg in ("C2", "C3", "C5"):
        whole = synth.make_states(cfg, 3000, start=1000)
        part = synth.make_states(cfg, 500, start=2500)
        for k in whole:
            assert np.array_equal(whole[k][..., 1500:2000], part[k])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,B,start", [("C1", 1, 0), ("C1", 37, 0), ("C2", 65536, 0), ("C3", 100003, 12345),
                                         ("C4", 4096, 1 << 20), ("C5", 70001, (1 << 24) - 70001), ("C5", 2048, 123456789)])
def test_device_generator_is_bit_identical(qlb_built, cfg, B, start):
    import torch
    from quadruped_locomotion_b200 import capi
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device")
    solver = capi.Solver("quadruped_model")
    dev = torch.device("cuda:0")
    ref = synth.make_states(cfg, B, start=start)
    for dt, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        d = {k: torch.full((n, B), -77.0, dtype=dt, device=dev) for k, n in (("q", 12), ("quat", 4), ("wrench", 6), ("mu", 4), ("normals", 12))}
        mask = torch.full((B,), 255, dtype=torch.uint8, device=dev)
        solver.generate_states(cfg, B, start=start, q=d["q"], quat=d["quat"], wrench=d["wrench"], mask=mask, mu=d["mu"],
                               normals=d["normals"], stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for k in d:
            got = d[k].cpu().numpy()
            want = ref[k].astype(npdt)
            assert np.array_equal(got.view(np.uint64 if npdt is np.float64 else np.uint32),
                                  want.view(np.uint64 if npdt is np.float64 else np.uint32)), (cfg, k, str(dt))
        assert np.array_equal(mask.cpu().numpy(), ref["mask"])
    # outputs are optional
    q = torch.zeros((12, B), dtype=torch.float64, device=dev)
    solver.generate_states(cfg, B, start=start, q=q)
    torch.cuda.synchronize()
    assert np.array_equal(q.cpu().numpy(), ref["q"])
    solver.close()
