"""Effect of the Newton tolerance on pass count and parity (phi+DM)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pp_oracle as orc
from tests import synth
from pulseportraiture_b200.engine import WidebandPlan

def run(nchan, nbin, seeds):
    cases = [synth.make_case(nchan, nbin, 1500., 800., s) for s in seeds]
    data = np.stack([c['data'] for c in cases]).astype(np.float32)
    P, freqs = cases[0]['P'], cases[0]['freqs']
    refs = []
    for c in cases:
        noise = orc.get_noise(c['data'], chans=True)
        ref, _, _ = orc.toa_core(c['data'], c['model'], P, freqs, noise, polish='exact')
        refs.append(ref)
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(cases[0]['model'].astype(np.float32), freqs)
        for tol in (1e-3, 1e-2, 3e-2, 1e-1):
            r = pl.fit_batch(data, P, tol=tol)
            n = len(cases)
            dchi = np.array([r['chi2'][i]/refs[i].chi2 - 1 for i in range(n)])
            dphi = np.array([(r['params'][i,0]-refs[i].phi)/refs[i].phi_err for i in range(n)])
            dDM = np.array([(r['params'][i,1]-refs[i].DM)/refs[i].DM_err for i in range(n)])
            derr = np.array([r['param_errs'][i,0]/refs[i].phi_err - 1 for i in range(n)])
            dsc = np.array([np.max(np.abs(r['scales'][i]/refs[i].scales - 1)) for i in range(n)])
            print('%4dx%-5d tol %.0e: passes %.2f | chi2 rel max %.1e | phi %.1e DM %.1e sigma | phi_err rel %.1e | scales rel %.1e'
                  % (nchan, nbin, tol, r['nfeval'].mean(), np.max(np.abs(dchi)), np.max(np.abs(dphi)), np.max(np.abs(dDM)), np.max(np.abs(derr)), np.max(dsc)))
run(64, 512, range(100, 132))
run(512, 2048, range(400, 408))
