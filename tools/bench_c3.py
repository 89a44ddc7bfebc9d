"""Config 3 (CHIME/uGMRT-like): 5-parameter scattering fit, 4096 chan x 1024 bin.
Prints TOAs/s and per-kernel times; not the headline metric (bench.py is config 2)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan

NCHAN, NBIN, NU0, BW = 4096, 1024, 600.0, 400.0
P = 1.0 / 345.67890123456789
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flags = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "11011")]
tau_s, alpha = 50e-6, -4.0
freqs = np.linspace(NU0 - BW / 2 + BW / (2.0 * NCHAN), NU0 + BW / 2 - BW / (2.0 * NCHAN), NCHAN)
gm = os.path.join(ROOT, "tests", "golden", "example.gmodel")
_, _, model = pplib.read_model(gm, pplib.get_bin_centers(NBIN), freqs, P, quiet=True)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
mFT = torch.fft.rfft(torch.from_numpy(model).to(dev), dim=-1)
k = torch.arange(mFT.shape[-1], device=dev, dtype=torch.float64)
taus = torch.from_numpy((tau_s / P) * (freqs / NU0) ** alpha).to(dev)
B = 1.0 / (1.0 + 2j * np.pi * taus[:, None] * k[None, :])
nu2 = torch.from_numpy(freqs ** -2.0 - NU0 ** -2.0).to(dev)
data = torch.empty((nsub, NCHAN, NBIN), dtype=torch.float32, device=dev)
phi = torch.rand(nsub, generator=g, device=dev, dtype=torch.float64) - 0.5
dDM = 3e-4 + 2e-4 * torch.randn(nsub, generator=g, device=dev, dtype=torch.float64)
for a in range(0, nsub, 16):
    b = min(nsub, a + 16)
    sh = -phi[a:b, None] - (pplib.Dconst * dDM[a:b, None] / P) * nu2[None, :]
    ph = torch.exp(2j * np.pi * (sh[:, :, None] * k[None, None, :]))
    clean = torch.fft.irfft(mFT[None] * B[None] * ph, n=NBIN, dim=-1)
    data[a:b] = clean.to(torch.float32) + 1.5 * torch.randn(clean.shape, generator=g, device=dev, dtype=torch.float32)
torch.cuda.synchronize()
pl = WidebandPlan(NCHAN, NBIN)
pl.set_model(np.ascontiguousarray(model, dtype=np.float64), freqs)
nu_fit = freqs.mean()
scat = np.tile([0.8 * (tau_s / P) * (nu_fit / NU0) ** alpha, alpha], (nsub, 1))
kw = dict(fit_flags=flags, log10_tau=True, scat_guess=scat, pinned_results=True)
fracs = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.0, 0.99]
base = None
for frac in fracs:
    pl.set_coarse(frac)
    pl.enable_timing(False)
    for _ in range(2):
        r = pl.fit_batch(data, P, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = pl.fit_batch(data, P, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    pl.enable_timing(True)
    r = pl.fit_batch(data, P, **kw)
    r = {k_: np.array(v) for k_, v in r.items() if isinstance(v, np.ndarray)}
    st = pl.stats()
    tau_out = 10 ** r["params"][:, 3] * (NU0 / r["nu_out"][:, 2]) ** r["params"][:, 4] * P
    line = {"workload": "config 3: 5-param fit %s, 4096x1024 x %d subints" % (flags, nsub), "coarse_frac": frac,
            "TOAs_per_s": nsub / dt, "ms_per_batch": dt * 1e3, "mean_evaluations": float(r["nfeval"].mean()),
            "converged": int((r["return_code"] == 0).sum()), "ms_spectra": st["ms_spectra"], "ms_pass": st["ms_pass"],
            "ms_update": st["ms_update"], "ms_coarse": st["ms_coarse"], "ms_guess": st["ms_guess"],
            "pass_launches": st["pass_launches"], "coarse_launches": st["coarse_launches"], "chunk": st["chunk"],
            "tau_at_600MHz_us_median": float(np.median(tau_out) * 1e6),
            "dDM_pull_rms": float(np.sqrt(np.mean(((r["params"][:, 1] - dDM.cpu().numpy()) / r["param_errs"][:, 1]) ** 2)))}
    if base is None:
        base = r
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.abs(r["params"] - base["params"]) / base["param_errs"]
        line["max_dparam_sigma_vs_first"] = float(np.nanmax(np.where(np.isfinite(d), d, 0.0)))
        line["max_rel_dchi2_vs_first"] = float(np.max(np.abs(r["chi2"] - base["chi2"]) / base["chi2"]))
        line["max_rel_derr_vs_first"] = float(np.nanmax(np.where(base["param_errs"] > 0, np.abs(r["param_errs"] / base["param_errs"] - 1), 0)))
    print(json.dumps(line))
