"""chi2 / parameter deviations vs the oracle for FFT precision x noise source."""
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
        refs.append((noise, ref))
    errs = np.stack([r[0] for r in refs])
    with WidebandPlan(nchan, nbin) as pl:
        pl.set_model(cases[0]['model'].astype(np.float32), freqs)
        for bits in (32, 64):
            pl.set_fft_precision(bits)
            for given in (False, True):
                r = pl.fit_batch(data, P, errs=errs if given else None)
                dchi = np.array([r['chi2'][i]/refs[i][1].chi2 - 1 for i in range(len(cases))])
                dphi = np.array([(r['params'][i,0]-refs[i][1].phi)/refs[i][1].phi_err for i in range(len(cases))])
                dDM = np.array([(r['params'][i,1]-refs[i][1].DM)/refs[i][1].DM_err for i in range(len(cases))])
                print('%4dx%-5d fft%d errs_%s: chi2 rel rms %.2e max %.2e | phi sig max %.2e DM sig max %.2e | passes %.2f'
                      % (nchan, nbin, bits, 'given' if given else 'meas ', np.sqrt(np.mean(dchi**2)), np.max(np.abs(dchi)),
                         np.max(np.abs(dphi)), np.max(np.abs(dDM)), r['nfeval'].mean()))

run(64, 512, range(100, 116))
run(16, 2048, range(200, 208))
run(256, 1024, range(300, 306))
run(512, 2048, range(400, 406))
