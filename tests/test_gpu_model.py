"""GPU tests of SURVEY 8(f) row f3: the evolving-Gaussian model portrait generated on
the device (pp_gen_gaussian_portrait) against the reference goldens and the oracle."""
import os

import numpy as np
import pytest

from oracle import pp_oracle as orc
from tests import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
GMODEL = os.path.join(HERE, "golden", "example.gmodel")
F32 = 1.0 / (1 << 23)   # float32 output: half an ulp relative to the profile maximum and then some


def _params(gm, tau_bin=0.0):
    return np.concatenate([[gm["dc"], tau_bin], np.asarray(gm["comps"], dtype=np.float64).ravel()])


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in G.files if k.startswith("model_")}))
def test_device_model_vs_reference_golden(case):
    from pulseportraiture_b200 import pplib
    nchan, nbin, nu0, bw = G[case + "/cfg"]
    nchan, nbin = int(nchan), int(nbin)
    freqs = orc.make_freqs(nchan, nu0, bw)
    _, ngauss, model = pplib.read_model(GMODEL, pplib.get_bin_centers(nbin), freqs, synth.P_EXAMPLE,
                                        quiet=True, device=True)
    assert ngauss == 3 and model.shape == (nchan, nbin)
    tol = F32 * np.max(np.abs(G[case + "/row0"]))
    assert np.max(np.abs(model[0] - G[case + "/row0"])) <= tol
    assert np.max(np.abs(model[-1] - G[case + "/rowlast"])) <= F32 * np.max(np.abs(G[case + "/rowlast"]))
    # identical to the host generator after rounding to float32, up to one float32 ulp
    _, _, host = pplib.read_model(GMODEL, pplib.get_bin_centers(nbin), freqs, synth.P_EXAMPLE, quiet=True)
    assert np.max(np.abs(model - host)) <= F32 * np.max(np.abs(host))


@pytest.mark.parametrize("code,tau_s,nchan,nbin,nu0,bw", [
    ("000", 0.0, 48, 256, 1500., 800.),
    ("000", 40e-6, 64, 1024, 600., 400.),     # scattered (TAU != 0): rfft * B, irfft
    ("110", 0.0, 32, 512, 1400., 600.),       # linear loc / wid evolution
    ("011", 15e-6, 20, 2048, 800., 200.),
    ("000", 5e-6, 8, 64, 1500., 800.),
])
def test_device_model_vs_oracle(code, tau_s, nchan, nbin, nu0, bw):
    from pulseportraiture_b200.engine import WidebandPlan
    gm = dict(orc.read_gmodel(GMODEL))
    gm["code"] = code
    comps = np.array(gm["comps"], dtype=np.float64)
    if code[0] == "1":
        comps[:, 1] = [2e-5, -1e-5, 3e-5]     # slopes [rot/MHz]
    if code[1] == "1":
        comps[:, 3] = [1e-5, -5e-6, 2e-6]
    if code[2] == "1":
        comps[:, 5] = [-2e-3, 1e-3, 5e-4]
    # one component across the phase wrap and a narrow one
    comps = np.vstack([comps, [0.985, comps[0, 1], 0.03, comps[0, 3], 1.5, comps[0, 5]],
                       [0.6, 0.0 if code[0] == "1" else 0.01, 0.004, 0.0 if code[1] == "1" else -0.5, 2.0,
                        0.0 if code[2] == "1" else -1.0]])
    gm["comps"] = comps
    gm["alpha"] = -3.7
    freqs = orc.make_freqs(nchan, nu0, bw)
    P = synth.P_EXAMPLE
    ref = orc.gen_gaussian_portrait(gm, orc.get_bin_centers(nbin), freqs, P=P, tau_override=tau_s)
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_freqs(freqs)
        out = pl.gen_gaussian_portrait(code, _params(gm, tau_s * nbin / P), gm["alpha"], gm["nu_ref"])
        assert out.dtype == np.float32
        assert np.max(np.abs(out - ref)) <= 2 * F32 * np.max(np.abs(ref))
        # the float64 generator (pp_gen_gaussian_portrait_f64): double evaluation when there is no scattering
        out64 = pl.gen_gaussian_portrait(code, _params(gm, tau_s * nbin / P), gm["alpha"], gm["nu_ref"], dtype=np.float64)
        assert out64.dtype == np.float64
        assert np.max(np.abs(out64 - ref)) <= (1e-13 if tau_s == 0 else 2 * F32) * np.max(np.abs(ref))
        pl.set_model(out64, freqs)                        # (float64 models go through pp_set_model_f64)
        assert 0.0 < pl.stats()["x_keep_frac"] <= 1.0
        # device-resident output feeds set_model directly: same fit as with the host model
        import torch
        dev = pl.gen_gaussian_portrait(code, _params(gm, tau_s * nbin / P), gm["alpha"], gm["nu_ref"], device_out=True)
        assert isinstance(dev, torch.Tensor) and dev.is_cuda
        assert np.array_equal(dev.cpu().numpy(), out)
        rng = np.random.RandomState(7)
        data = orc.rotate_data(ref, -0.07, -2e-4, P, freqs, nu0) + 0.3 * rng.standard_normal(ref.shape)
        data = data.astype(np.float32)
        pl.set_model(dev, freqs)
        r1 = pl.fit_batch(data[None], P)
        pl.set_model(out, freqs)
        r2 = pl.fit_batch(data[None], P)
    assert np.array_equal(r1["params"], r2["params"]) and np.array_equal(r1["chi2"], r2["chi2"])
    assert abs(r1["params"][0, 1] - 2e-4) < 6 * r1["param_errs"][0, 1]


def test_device_model_argument_errors():
    from pulseportraiture_b200.engine import WidebandPlan
    from pulseportraiture_b200._ffi import PPError
    gm = orc.read_gmodel(GMODEL)
    with WidebandPlan(16, 128) as pl:
        with pytest.raises(PPError):     # frequencies not set
            pl.gen_gaussian_portrait("000", _params(gm), -4.0, gm["nu_ref"])
        pl.set_freqs(orc.make_freqs(16, 1500., 800.))
        with pytest.raises(PPError):
            pl.gen_gaussian_portrait("0x0", _params(gm), -4.0, gm["nu_ref"])
        with pytest.raises(ValueError):
            pl.gen_gaussian_portrait("000", np.zeros(9), -4.0, gm["nu_ref"])
        # no components: DC only
        out = pl.gen_gaussian_portrait("000", np.array([0.25, 0.0]), -4.0, 1400.0)
        assert np.all(out == np.float32(0.25))


def _spline_model(nbin, ncomp, k, s, seed):
    """A synthetic make_spline_model-style model: mean profile, orthonormal eigenvectors and a
    B-spline through noisy projections (scipy.interpolate.splprep, as pplib.py:1203-1206)."""
    import scipy.interpolate as si
    rng = np.random.RandomState(seed)
    x = orc.get_bin_centers(nbin)
    mean_prof = np.exp(-0.5 * ((x - 0.3) / 0.02) ** 2) + 0.3 * np.exp(-0.5 * ((x - 0.36) / 0.05) ** 2)
    eigvec = np.linalg.qr(rng.standard_normal((nbin, max(ncomp, 1))))[0][:, :ncomp]
    if not ncomp:
        return mean_prof, eigvec, (np.zeros(8), [], 3)
    f = np.linspace(1150., 1850., 48)
    proj = np.array([0.2 * np.sin(f / (120. + 40 * i) + i) + 0.01 * rng.standard_normal(len(f)) for i in range(ncomp)])
    tck, _ = si.splprep(proj, u=f, k=k, s=s)
    return mean_prof, eigvec, tck


@pytest.mark.parametrize("nchan,nbin,ncomp,k,s", [(64, 512, 4, 3, 0.02), (33, 2048, 10, 3, 0.0), (16, 256, 2, 5, 0.05),
                                                   (8, 128, 1, 1, 0.1), (12, 1024, 0, 3, 0.0)])
def test_device_spline_model_vs_scipy(nchan, nbin, ncomp, k, s, tmp_path):
    import pickle
    from pulseportraiture_b200 import pplib
    mean_prof, eigvec, tck = _spline_model(nbin, ncomp, k, s, 11 * nbin + ncomp)
    freqs = orc.make_freqs(nchan, 1500., 800.)      # 1100-1900 MHz: extrapolates beyond the knots
    ref = pplib.gen_spline_portrait(mean_prof, freqs, eigvec, tck)          # scipy splev + dot, as the reference
    out = pplib.gen_spline_portrait(mean_prof, freqs, eigvec, tck, device=True)
    assert out.shape == (nchan, nbin)
    assert np.max(np.abs(out - ref)) <= 2 * F32 * np.max(np.abs(ref))
    # through the pickle reader, as GetTOAs finds its model
    path = str(tmp_path / "model.spl")
    with open(path, "wb") as fh:
        pickle.dump(["mdl", "J0000+0000", "none", mean_prof, eigvec, tck], fh, protocol=2)
    assert pplib.is_spline_model(path) and not pplib.is_spline_model(GMODEL)
    name, model = pplib.read_spline_model(path, freqs, nbin, quiet=True, device=True)
    assert name == "mdl" and np.array_equal(model, out)


@pytest.mark.parametrize("nbin_model,nbin,ncomp", [(512, 1024, 3), (1024, 256, 4), (256, 2048, 0)])
def test_spline_model_resampled_to_another_nbin(nbin_model, nbin, ncomp):
    """gen_spline_portrait(..., nbin != len(mean_prof)) (pplib.py:951-955): ss.resample along the bin
    axis, then the rotation by half the bin-width difference.  Restated here per channel, as the
    reference does it; the product resamples the basis instead (linear), on host or device."""
    import scipy.interpolate as si
    import scipy.signal as ss
    from pulseportraiture_b200 import pplib
    mean_prof, eigvec, tck = _spline_model(nbin_model, ncomp, 3, 0.02, 5 * nbin + ncomp)
    freqs = orc.make_freqs(24, 1500., 700.)
    if ncomp:
        port = np.dot(np.array(si.splev(freqs, tck, der=0, ext=0)).T, eigvec.T) + mean_prof
    else:
        port = np.tile(mean_prof, (len(freqs), 1))
    shift = 0.5 * (nbin ** -1.0 - nbin_model ** -1.0)
    ref = orc.rotate_data(ss.resample(port, nbin, axis=1), shift)
    for device in (False, True):
        out = pplib.gen_spline_portrait(mean_prof, freqs, eigvec, tck, nbin=nbin, device=device)
        assert out.shape == (len(freqs), nbin)
        assert np.max(np.abs(out - ref)) <= 4 * F32 * np.max(np.abs(ref))
