"""Config 5: the ppalign / ppzap work on 2000 subints of 512 x 2048 with the
data resident on the device: niter = 3 of {FFTFIT guess with Ns = nbin -> phi+DM fit ->
Fourier-domain rotate -> weighted accumulate}, then a per-channel fit_phase_shift scan
(nsub x nchan profiles against the mean profile, Ns = 100) and the ppzap thresholds.
Prints one JSON line; not the headline metric."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pulseportraiture_b200 import pplib, ppzap
from pulseportraiture_b200.engine import WidebandPlan

NCHAN, NBIN, NU0, BW = 512, 2048, 1500.0, 800.0
P = 1.0 / 345.67890123456789
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
freqs = np.linspace(NU0 - BW / 2 + BW / (2.0 * NCHAN), NU0 + BW / 2 - BW / (2.0 * NCHAN), NCHAN)
gm = os.path.join(ROOT, "tests", "golden", "example.gmodel")
_, _, model = pplib.read_model(gm, pplib.get_bin_centers(NBIN), freqs, P, quiet=True)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(55)
mFT = torch.fft.rfft(torch.from_numpy(model).to(dev), dim=-1)
k = torch.arange(mFT.shape[-1], device=dev, dtype=torch.float64)
nu2 = torch.from_numpy(freqs ** -2.0 - NU0 ** -2.0).to(dev)
data = torch.empty((nsub, NCHAN, NBIN), dtype=torch.float32, device=dev)
phi = torch.rand(nsub, generator=g, device=dev, dtype=torch.float64) - 0.5
dDM = 3e-4 + 2e-4 * torch.randn(nsub, generator=g, device=dev, dtype=torch.float64)
for a in range(0, nsub, 64):
    b = min(nsub, a + 64)
    sh = -phi[a:b, None] - (pplib.Dconst * dDM[a:b, None] / P) * nu2[None, :]
    ph = torch.exp(2j * np.pi * (sh[:, :, None] * k[None, None, :]))
    clean = torch.fft.irfft(mFT[None] * ph, n=NBIN, dim=-1)
    data[a:b] = clean.to(torch.float32) + 1.5 * torch.randn(clean.shape, generator=g, device=dev, dtype=torch.float32)
torch.cuda.synchronize()

pl = WidebandPlan(NCHAN, NBIN)
# start template: a smoothed, deliberately mis-aligned copy of the model
template = np.roll(model, 37, axis=1)
times = {"fused": [], "fit": [], "accumulate": []}
t_all0 = time.perf_counter()
for it in range(3):
    pl.set_model(template.astype(np.float32), freqs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = pl.fit_batch(data, P, nu_fit_mode=1, Ns=NBIN, pinned_results=True, align=True)   # fit + rotate + accumulate
    torch.cuda.synchronize(); t1 = time.perf_counter()
    template = r["align_sum"] / r["align_wsum"][:, None]
    times["fused"].append(t1 - t0)
t_align = time.perf_counter() - t_all0
# the two-step form of the last iteration, for comparison: fit, then pp_align_accumulate
pl.set_model(np.roll(model, 37, axis=1).astype(np.float32), freqs)
torch.cuda.synchronize(); t0 = time.perf_counter()
r2 = pl.fit_batch(data, P, nu_fit_mode=1, Ns=NBIN, pinned_results=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
w = r2["scales"] / np.where(r2["noise"] > 0, r2["noise"], 1.0) ** 2
acc, wsum = pl.align_accumulate(data, r2["params"][:, 0].copy(), r2["params"][:, 1].copy(), P, r2["nu_out"][:, 0].copy(), np.ascontiguousarray(w))
torch.cuda.synchronize(); t2 = time.perf_counter()
times["fit"].append(t1 - t0); times["accumulate"].append(t2 - t1)
# alignment quality: the template converges to a rotated model
g0 = pplib.fit_phase_shift(template.mean(0), model.mean(0), Ns=NBIN)
# per-channel FFTFIT scan against the mean profile (pplib.py:2497 / pptoas.py:992) and zap thresholds
prof = template.mean(axis=0).astype(np.float32)
torch.cuda.synchronize(); t0 = time.perf_counter()
ps = pl.fit_phase_shift_batch(data.view(nsub * NCHAN, NBIN), prof[None], Ns=100)
torch.cuda.synchronize(); t_scan = time.perf_counter() - t0
t0 = time.perf_counter()
noise_stds = pl.get_noise_batch(data)
d = pplib.DataBunch(noise_stds=noise_stds[:, None, :], ok_isubs=np.arange(nsub), ok_ichans=[np.arange(NCHAN)] * nsub,
                    nsub=nsub, nchan=NCHAN)
zap = ppzap.get_zap_channels(d, nstd=3)
t_zap = time.perf_counter() - t0
print(json.dumps({"workload": "config 5: ppalign niter=3 + per-channel FFTFIT scan + zap thresholds, %d subints of 512x2048 (device-resident)" % nsub,
                  "fused_align_s_per_iteration": [round(x, 4) for x in times["fused"]],
                  "two_step_fit_s": round(times["fit"][0], 4), "two_step_accumulate_s": round(times["accumulate"][0], 4),
                  "align_total_s": round(t_align, 3), "subint_iterations_per_s": round(3 * nsub / t_align, 1),
                  "template_vs_model_phase": float(g0.phase), "template_vs_model_snr": float(g0.snr),
                  "scan_profiles": nsub * NCHAN, "scan_s": round(t_scan, 3), "scan_profiles_per_s": round(nsub * NCHAN / t_scan, 1),
                  "zap_s": round(t_zap, 3), "zapped": int(sum(len(z) for z in zap))}))
