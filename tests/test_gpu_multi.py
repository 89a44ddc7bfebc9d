"""The product's own multi-GPU split on hardware: one batch sharded in contiguous
ranges over several plans (one host thread, plan and stream per device) must give the arrays a single
plan gives, bit for bit; the fused ppalign sums of the shards add up."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _batch(n, nchan=64, nbin=512, seed0=500):
    cases = [synth.make_case(nchan, nbin, 1500., 800., seed0 + i) for i in range(n)]
    data = np.stack([c["data"] for c in cases]).astype(np.float32)
    return cases[0], data


def _devices(want):
    import torch
    nd = torch.cuda.device_count()
    return [i % nd for i in range(want)], nd


@pytest.mark.parametrize("world", [2, 3])
def test_multigpu_fitter_matches_one_plan_bit_for_bit(world):
    from pulseportraiture_b200.engine import WidebandPlan
    from pulseportraiture_b200.multigpu import MultiGPUFitter
    c, data = _batch(11)
    devs, nd = _devices(world)          # two different GPUs where the box has them, else plans on one GPU
    errs = np.full((11, 64), 1.5)
    errs[:, 5] = 2.0
    mask = np.ones((11, 64), dtype=np.uint8)
    mask[3, 10:20] = 0
    kw = dict(errs=errs, chan_mask=mask, DM_guess=np.linspace(0, 1e-4, 11), fit_flags=(1, 1, 0, 0, 0))
    with WidebandPlan(64, 512, device=0) as pl:
        pl.set_model(c["model"].astype(np.float32), c["freqs"])
        one = pl.fit_batch(data, c["P"], **kw)
    mf = MultiGPUFitter(64, 512, devs)
    try:
        mf.set_model(c["model"].astype(np.float32), c["freqs"])
        many = mf.fit_batch(data, c["P"], **kw)
    finally:
        mf.close()
    assert set(many) == set(one)
    for k in one:
        assert many[k].shape == one[k].shape, k
        assert np.array_equal(many[k], one[k], equal_nan=True), k
    assert np.all(one["return_code"] == 0)


def test_multigpu_align_sums_add_up():
    from pulseportraiture_b200.engine import WidebandPlan
    from pulseportraiture_b200.multigpu import MultiGPUFitter
    c, data = _batch(8, seed0=700)
    devs, _ = _devices(2)
    kw = dict(fit_flags=(1, 1, 0, 0, 0), Ns=512, align=True, nu_fit_mode=1)
    with WidebandPlan(64, 512, device=0) as pl:
        pl.set_model(c["model"].astype(np.float32), c["freqs"])
        one = pl.fit_batch(data, c["P"], **kw)
    mf = MultiGPUFitter(64, 512, devs)
    try:
        mf.set_model(c["model"].astype(np.float32), c["freqs"])
        many = mf.fit_batch(data, c["P"], **kw)
    finally:
        mf.close()
    assert many["align_sum"].shape == (64, 512) and many["align_wsum"].shape == (64,)
    assert np.array_equal(many["params"], one["params"])
    scale = np.abs(one["align_sum"]).max()
    assert np.abs(many["align_sum"] - one["align_sum"]).max() < 1e-12 * scale     # summation order differs
    assert np.allclose(many["align_wsum"], one["align_wsum"], rtol=1e-13, atol=0)


def test_fit_args_that_look_per_subint_are_not_sliced():
    """fit_flags has five entries: a batch of five subints must not have it cut per shard."""
    from pulseportraiture_b200.multigpu import MultiGPUFitter
    c, data = _batch(5, seed0=900)
    devs, _ = _devices(2)
    mf = MultiGPUFitter(64, 512, devs)
    try:
        mf.set_model(c["model"].astype(np.float32), c["freqs"])
        r = mf.fit_batch(data, c["P"], fit_flags=np.array([1, 1, 0, 0, 0]),
                         bounds=[(None, None), (-1.0, 1.0), None, None, None])
    finally:
        mf.close()
    assert r["params"].shape == (5, 5) and np.all(r["return_code"] == 0)
