"""Array-of-structs face of the wrench-mode solve (qlb_solve_records[_host]) against the SoA entry point:
bit-identical results, ragged sizes, several pipeline chunks."""
import numpy as np
import pytest

from quadruped_locomotion_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 7, 130, 4097, 300001])
def test_records_host_equals_soa_host(qlb_built, B):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device")
    solver = capi.Solver("quadruped_model")
    st = synth.make_states("C5" if B > 1 else "C1", B, start=31)
    st["normals"] = None                      # the record entry uses the default normals
    ref = solver.solve_wrench_numpy(st)
    rec = capi.wrench_records(st)
    out = np.zeros(B, dtype=capi.RESULT_RECORD_DTYPE)
    out["grf"] = 7.0
    solver.solve_records_host(rec, out)
    assert np.array_equal(out["grf"].T, ref["grf"]) and np.array_equal(out["tau"].T, ref["tau"])
    assert np.array_equal(out["netwrench"].T, ref["netwrench"]) and np.array_equal(out["flags"], ref["flags"])
    assert (out["reserved"] == 0).all()
    # device-pointer twin
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros(B * capi.RESULT_RECORD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    solver.solve_records(d_rec, d_out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(capi.RESULT_RECORD_DTYPE)
    assert np.array_equal(got["grf"], out["grf"]) and np.array_equal(got["flags"], out["flags"])
    solver.close()
