"""How the (phi, DM) device solver and the oracle's TNC behave from start values far from the
optimum (fit_portrait with a caller-supplied init_params instead of the FFTFIT guess)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pp_oracle as orc
from tests import synth
from pulseportraiture_b200.engine import WidebandPlan

c = synth.make_case(64, 512, 1500., 800., 4242, phi=0.123, dDM=3e-4)
data = c["data"].astype(np.float32)[None]
with WidebandPlan(64, 512) as pl:
    pl.set_model(c["model"].astype(np.float32), c["freqs"])
    for phi0 in (0.125, 0.135, 0.15, 0.17, 0.2, 0.05):
        ref = orc.fit_portrait(c["data"], c["model"], [phi0, 0.0], c["P"], c["freqs"])
        for mi in (8, 40):
            init = np.array([[phi0, 0, 0, 0, 0.]])
            r = pl.fit_batch(data, c["P"], init=init, max_iter=mi, semantics="fit_portrait")
            print("phi0 %.3f max_iter %2d: rc %d nfev %2d phase %.6f DM %.3e chi2 %.2f | TNC rc %d nfev %d phase %.6f DM %.3e chi2 %.2f"
                  % (phi0, mi, r["return_code"][0], r["nfeval"][0], r["params"][0, 0], r["params"][0, 1], r["chi2"][0],
                     ref.return_code, ref.nfeval, ref.phase, ref.DM, ref.chi2))
