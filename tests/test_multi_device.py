"""The single-process multi-GPU entry (mgb200_multi_*, SURVEY.md 8(b) "Threading"): ONE handle, global arrays in, the
library slices the z-slabs and drives every device on its own host thread.  With one device it must reproduce the plain
handle bit for bit; with two or more (skipped on a one-GPU box) the row-partitioned run must match the GLOBAL CPU oracle."""
import numpy as np
import pytest

from conftest import make_problem

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_multi_handle_on_one_device_equals_the_plain_handle():
    import multigrid_jl_b200 as mg
    A, AT, M, p, b = make_problem("poisson", [24, 20, 16], 3, maxit=6)
    dev = mg.DeviceHierarchy(p, device=0)
    dev.set_cycle(p)
    x0, it0, res0 = dev.solveMG(b, np.zeros_like(b), 1e-12, 6)
    xk0, itk0, flag0, resk0 = dev.solveCG(b, np.zeros_like(b), 1e-8, 30)
    dev.destroy()
    md = mg.MultiDeviceHierarchy(p, [0], index_base=1)
    x1, it1, res1 = md.solveMG(b, np.zeros_like(b), 1e-12, 6)
    xk1, itk1, flag1, resk1 = md.solveCG(b, np.zeros_like(b), 1e-8, 30)
    z = md.precondition(b)
    md.destroy()
    assert it0 == it1 and np.array_equal(res0, res1) and np.array_equal(x0, x1)
    assert itk0 == itk1 and flag0 == flag1 and np.array_equal(resk0, resk1) and np.array_equal(xk0, xk1)
    assert np.isfinite(z).all()


@pytest.mark.parametrize("kind,n,cycle", [("poisson", [32, 32, 64], 'V'), ("helmholtz", [24, 24, 48], 'W')])
def test_multi_handle_two_devices_matches_the_global_oracle(kind, n, cycle):
    G = min(_n_gpus(), 2 if n[2] < 96 else 4)
    if G < 2:
        pytest.skip("needs two GPUs")
    import multigrid_jl_b200 as mg
    from oracle import cycle as oc
    A, AT, M, p, b = make_problem(kind, n, 4, cycle=cycle, maxit=5)
    o = oc.OracleMG(p)
    x_ref, it_ref, res_ref = oc.solveMG(o, b, np.zeros_like(b))
    md = mg.MultiDeviceHierarchy(p, list(range(G)), replicate_below=3000)
    info = md.info()
    assert info["world"] == G and info["dist_levels"] >= 1
    x, it, res = md.solveMG(b, np.zeros_like(b), p.relativeTol, p.maxOuterIter)
    assert it == it_ref
    np.testing.assert_allclose(res, res_ref, rtol=1e-10)
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    if kind == "poisson":
        p.maxOuterIter, p.relativeTol = 30, 1e-8
        o = oc.OracleMG(p)
        xr, itc_ref, flag_ref, resc_ref = oc.solveCG_MG(AT, o, b, np.zeros_like(b))
        xc, itc, flag, resc = md.solveCG(b, np.zeros_like(b), 1e-8, 30)
        assert itc == itc_ref and flag == flag_ref == 0
        np.testing.assert_allclose(resc, resc_ref[:itc], rtol=1e-8)
    md.destroy()
