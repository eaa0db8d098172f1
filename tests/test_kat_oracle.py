"""Known-answer tests against the CPU oracles."""
import numpy as np
import pytest
import torch

from oracle import magnet_oracle as mo
from oracle import torch_ref
from tests import kat
from tests.util import oracle_cw


@pytest.mark.parametrize("case", sorted(kat.CW_CASES))
def test_cw_known_answers(case):
    inp, dvol, exp, tol = kat.CW_CASES[case]()
    got = oracle_cw(inp, dvol.numpy())
    port = torch_ref.cost_volume_cw(dvol, inp.ref_feat, inp.nghbr_feat, inp.ref_gmms, inp.nghbr_gmms, inp.R, inp.t,
                                    inp.is_valid, inp.cam_intrins, inp.thres).numpy()
    for name, out in (("numpy", got), ("aten", port)):
        if tol == 0.0:
            assert np.array_equal(out, exp.astype(np.float32)), (case, name)
        else:
            assert np.abs(out - exp).max() <= tol * max(np.abs(exp).max(), 1.0), (case, name, np.abs(out - exp).max())
    if case == "one_pixel_shift":
        assert np.array_equal(got[..., -1], np.zeros_like(got[..., -1]))     # zero padding, exactly


def test_f_identity_uniform():
    inp, planes, exp, tol = kat.f_identity()
    got = mo.cost_volume_f(planes, inp.ref_feat.numpy(), inp.nghbr_feat.numpy(), inp.R.numpy(), inp.t.numpy(),
                           inp.is_valid.numpy(), inp.cam_intrins['intM'].numpy(),
                           inp.cam_intrins['unit_ray_array_2D'].numpy())
    assert np.abs(got - exp).max() <= 1e-6
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-6)
