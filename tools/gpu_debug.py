"""First-light diagnostics on the GPU box: dumps intermediate quantities next
to the oracle's.  Not a test; see tests/test_gpu_*.py."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pp_oracle as orc
from tests import synth
from pulseportraiture_b200.engine import WidebandPlan

def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

def one(nchan, nbin, seed, nu0=1500., bw=800., **kw):
    c = synth.make_case(nchan, nbin, nu0, bw, seed, **kw)
    data, model, freqs, P = c['data'], c['model'], c['freqs'], c['P']
    pl = WidebandPlan(nchan, nbin)
    pl.enable_timing(True)
    pl.set_model(model.astype(np.float32), freqs)
    t = time.time()
    r = pl.fit_batch(data.astype(np.float32)[None], P, want_chan_sums=True)
    dt = time.time() - t
    # oracle
    noise = orc.get_noise(data, chans=True)
    res, phi_guess, nu_fits = orc.toa_core(data, model, P, freqs, noise, polish='exact')
    print('--- %dx%d seed %d  (%.1f ms) stats %s' % (nchan, nbin, seed, dt*1e3, pl.stats()))
    print(' noise rel', rel(r['noise'][0], noise))
    print(' lag', r['lag_index'][0], res.lag_index, ' phi_guess', r['phi_guess'][0], phi_guess)
    print(' phi  %.12f vs %.12f  d/sig %.3e' % (r['params'][0,0], res.phi, (r['params'][0,0]-res.phi)/res.phi_err))
    print(' DM   %.6e vs %.6e  d/sig %.3e' % (r['params'][0,1], res.DM, (r['params'][0,1]-res.DM)/res.DM_err))
    print(' errs rel', rel(r['param_errs'][0,:2], [res.phi_err, res.DM_err]), ' nu_DM', r['nu_out'][0,0], res.nu_DM)
    print(' chi2 %.6f vs %.6f rel %.3e   red %.8f' % (r['chi2'][0], res.chi2, r['chi2'][0]/res.chi2-1, r['red_chi2'][0]))
    print(' snr rel', rel(r['snr'][0], res.snr), ' scales rel', rel(r['scales'][0], res.scales),
          ' scale_errs rel', rel(r['scale_errs'][0], res.scale_errs), ' csnr rel', rel(r['channel_snrs'][0], res.channel_snrs))
    print(' cov', r['cov'][0,:2,:2].ravel(), np.asarray(res.covariance_matrix).ravel())
    print(' nfeval', r['nfeval'][0], 'rc', r['return_code'][0], ' oracle nfev', res.nfeval)
    pl.close()

if __name__ == '__main__':
    one(64, 512, 0, phi=0.123, dDM=3e-4, legacy_seed=True)
    one(64, 512, 103)
    one(32, 256, 101)
    one(16, 1024, 104)
    one(48, 128, 105, nu0=430., bw=100.)
    one(8, 2048, 106)
    one(8, 4096, 107)
    one(8, 64, 108)
    one(512, 2048, 7)
