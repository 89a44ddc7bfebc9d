"""Coarse-to-fine start of the general solver at low S/N and with poor start values: return codes, evaluations
and agreement with the plain iterations (coarse_frac = 0)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan

nchan, nbin, nu0, bw, nsub = 256, 1024, 600.0, 400.0, 256
tau_s, alpha = 50e-6, -4.0
P = bench.P_EXAMPLE
freqs = np.linspace(nu0 - bw / 2 + bw / (2.0 * nchan), nu0 + bw / 2 - bw / (2.0 * nchan), nchan)
_, _, model = pplib.read_model(bench.GMODEL, pplib.get_bin_centers(nbin), freqs, P, quiet=True)
dev = torch.device("cuda", 0)
phi, dDM = bench.global_draws(nsub, 77)
out = []
for sigma in (1.5, 6.0, 20.0, 60.0):
    for tau_start in (0.8, 0.3, 3.0):
        bench.SIGMA, old = sigma, bench.SIGMA
        data = bench.make_device_batch(model, freqs, phi, dDM, 77, dev, nchan, nbin, nu0, scatter=(tau_s / P, alpha))
        bench.SIGMA = old
        scat = np.tile([tau_start * (tau_s / P) * (freqs.mean() / nu0) ** alpha, alpha], (nsub, 1))
        res = {}
        for frac in (0.0, 0.99):
            with WidebandPlan(nchan, nbin) as pl:
                pl.set_model(model.astype(np.float32), freqs)
                pl.set_coarse(frac)
                for flags in ((1, 1, 0, 1, 1), (1, 1, 1, 1, 1)):
                    r = pl.fit_batch(data, P, fit_flags=flags, log10_tau=True, scat_guess=scat)
                    st = pl.stats()
                    res[(frac, flags)] = ({k: np.array(v) for k, v in r.items()}, st)
        for flags in ((1, 1, 0, 1, 1), (1, 1, 1, 1, 1)):
            (a, sa), (b, sb) = res[(0.0, flags)], res[(0.99, flags)]
            both = (a["return_code"] == 0) & (b["return_code"] == 0)
            fit = np.array(flags, bool)
            with np.errstate(invalid="ignore", divide="ignore"):
                d = np.abs(a["params"] - b["params"])[:, fit] / a["param_errs"][:, fit]
            out.append({"sigma": sigma, "tau_start": tau_start, "flags": "".join(map(str, flags)),
                        "rc0_plain": int((a["return_code"] == 0).sum()), "rc0_coarse": int((b["return_code"] == 0).sum()),
                        "eval_plain": float(a["nfeval"].mean()), "eval_coarse": float(b["nfeval"].mean()),
                        "full_launches_plain": int(sa["pass_launches"]), "full_launches_coarse": int(sb["pass_launches"]),
                        "coarse_launches": int(sb["coarse_launches"]), "snr_median": float(np.median(a["snr"])),
                        "chi2_coarse_minus_plain_where_differ": [float(v) for v in (b["chi2"] - a["chi2"])[both][np.nanmax(d[both], axis=1) > 1e-2][:8]],
                        "max_dparam_sigma_both_converged": float(np.nanmax(d[both])) if both.any() else None,
                        "frac_within_1e-3": float(np.mean(np.nanmax(d[both], axis=1) < 1e-3)) if both.any() else None})
        del data
for o in out:
    print(json.dumps(o))
