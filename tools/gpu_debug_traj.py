"""Trajectory of the general (five-parameter) device solver on a golden_v2 case: the state after
1, 2, ... passes (max_iter = k), against the reference's converged values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import synth
from pulseportraiture_b200.engine import WidebandPlan
G2 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))
case = sys.argv[1] if len(sys.argv) > 1 else "full15_314"
cfg = G2[case + "/cfg"]
nchan, nbin, nu0, bw, seed = int(cfg[0]), int(cfg[1]), cfg[2], cfg[3], int(cfg[4])
tau_s, log10, option, sigma = cfg[5], bool(cfg[6]), int(cfg[7]), cfg[8]
flags = [int(v) for v in G2[case + "/flags"]]
c = synth.make_case(nchan, nbin, nu0, bw, seed, tau_data_s=tau_s, sigma=sigma)
g = lambda f: G2[case + "/full." + f]
print(case, flags, "log10" if log10 else "lin", "ref params", g("params"), "errs", g("param_errs"), "chi2", g("chi2"))
with WidebandPlan(nchan, nbin) as pl:
    pl.set_model(c["model"].astype(np.float32), c["freqs"])
    for k in list(range(1, 16)) + [20, 30, 40, 60]:
        r = pl.fit_batch(c["data"].astype(np.float32)[None], c["P"], errs=G2[case + "/errs"][None],
                         init=np.array(G2[case + "/init"], dtype=np.float64)[None], fit_flags=flags,
                         log10_tau=log10, option=option, max_iter=k)
        d = (r["params"][0] - g("params")) / np.where(g("param_errs") > 0, g("param_errs"), 1)
        print("max_iter %2d rc %d nfev %2d chi2-ref %+.3e  dev/sigma %s  nu_out %s" % (
            k, r["return_code"][0], r["nfeval"][0], r["chi2"][0] - g("chi2"), np.array2string(d, precision=3), r["nu_out"][0]))
