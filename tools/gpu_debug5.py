"""5-parameter path diagnostics vs the reference goldens and the oracle primitives."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pp_oracle as orc
from tests import synth
from pulseportraiture_b200.engine import WidebandPlan
G = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
cases = sorted({k.split("/")[0] for k in G.files if k.startswith("full_")})
def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
for case in cases:
    cfg = G[case + "/cfg"]
    nchan, nbin, nu0, bw, seed = int(cfg[0]), int(cfg[1]), cfg[2], cfg[3], int(cfg[4])
    tau_s, log10, option = cfg[5], bool(cfg[6]), int(cfg[7])
    flags = [int(v) for v in G[case + "/flags"]]
    c = synth.make_case(nchan, nbin, nu0, bw, seed, tau_data_s=tau_s, sigma=0.5)
    errs = G[case + "/errs"]; init = np.array(G[case + "/init"], dtype=np.float64)[None]
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(c["model"].astype(np.float32), c["freqs"])
        r = pl.fit_batch(c["data"].astype(np.float32)[None], c["P"], errs=errs[None], init=init,
                         fit_flags=flags, log10_tau=log10, option=option, want_chan_sums=True)
    g = lambda f: G[case + "/full." + f]
    names = ["phi", "DM", "GM", "tau", "alpha"]
    dev = ["%s %.1e" % (nm, abs(r["params"][0, i] - g(nm)) / g(nm + "_err")) for i, nm in enumerate(names) if flags[i]]
    # oracle primitives at the GPU's final point, fit reference = mean freq; final params are at nu_out -> use oracle re-eval at nu_out
    dFT, mFT = orc._spectra(c["data"], c["model"])
    eF = errs * np.sqrt(nbin / 2.0)
    no = r["nu_out"][0]
    prob = orc._FullProblem(dFT, mFT, eF, c["P"], c["freqs"], no[0], no[1], no[2], flags, log10)
    pr = prob.primitives(list(r["params"][0]), order=2)
    cs = r["chan_sums"][0]
    keys = ["C", "Cth", "Cthth", "Ct", "Ctt", "Ctht", "S", "St", "Stt"]
    pd = []
    for i, k in enumerate(keys):
        ref = pr[k]; sc = np.max(np.abs(ref)) or 1.0
        pd.append("%s %.1e" % (k, np.max(np.abs(cs[:, i] - ref)) / sc))
    fo = -(pr["C"] ** 2 / pr["S"]).sum()
    fg = -(cs[:, 0] ** 2 / cs[:, 6]).sum()
    Sd = ((np.abs(dFT) ** 2) / eF[:, None] ** 2).sum()
    print(case, flags, "log10" if log10 else "lin", "rc", r["return_code"][0], "nfev", r["nfeval"][0], "ref nfev", int(g("nfeval")))
    print("   params[sig]:", ", ".join(dev), "| chi2 rel %.2e" % (r["chi2"][0] / g("chi2") - 1),
          "| f(gpu sums) vs f(oracle prims at same x) rel %.2e" % (fg / fo - 1), "| chi2gpu-(Sd+fg) %.2e" % (r["chi2"][0] - (Sd + fg)))
    print("   nu_out rel %.1e errs rel %.1e scales %.1e scale_errs %.1e" % (
        rel(r["nu_out"][0], [g("nu_DM"), g("nu_GM"), g("nu_tau")]),
        rel(r["param_errs"][0][np.array(flags, bool)], g("param_errs")[np.array(flags, bool)]),
        rel(r["scales"][0], g("scales")), rel(r["scale_errs"][0], g("scale_errs"))))
    print("   prims:", ", ".join(pd))
