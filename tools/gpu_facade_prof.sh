#!/bin/bash
# cProfile of GetTOAs.get_TOAs (int16 and float64 archives, 256 subints of 512 x 2048): where the host time goes
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r02_facade_prof.txt
import sys, time, json, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from pulseportraiture_b200 import pptoas
from pulseportraiture_b200.pplib import DataBunch, get_bin_centers
freqs, model = bench.make_model()
dev = torch.device('cuda', 0)
nsub = 256
phi, dDM = bench.global_draws(nsub)
data = bench.make_device_batch(model, freqs, phi, dDM, 1, dev)
sub32 = data[:nsub].cpu().numpy()
raw, scl, offs, dec = pptoas.quantize_subints(sub32)
NCHAN, NBIN = bench.NCHAN, bench.NBIN
common = dict(backend="GUPPI", backend_delay=0.0, bw=bench.BW, doppler_factors=np.ones(nsub), DM=0.0, dmc=0,
              epochs=[pptoas.MJD(56000 + i // 100, 0.01 * (i % 100)) for i in range(nsub)], filename="bench.npz",
              freqs=np.tile(freqs, (nsub, 1)), frontend="Rcvr_800", integration_length=float(nsub), masks=None,
              nbin=NBIN, nchan=NCHAN, npol=1, nsub=nsub, nu0=bench.NU0, ok_ichans=[np.arange(NCHAN)] * nsub,
              ok_isubs=np.arange(nsub), parallactic_angles=np.zeros(nsub), phases=get_bin_centers(NBIN),
              Ps=np.full(nsub, bench.P_EXAMPLE), SNRs=np.ones((nsub, 1, NCHAN)), source="J0000+0000", state="Intensity",
              subtimes=np.ones(nsub), telescope="GBT", telescope_code="1", weights=np.ones((nsub, NCHAN)),
              noise_stds=np.full((nsub, 1, NCHAN), bench.SIGMA), flux_prof=None, prof=None, prof_noise=None, prof_SNR=None)
for name, extra in (("i16", dict(subints=dec[:, None], raw_subints=raw, dat_scl=scl, dat_offs=offs)),
                    ("f64", dict(subints=sub32[:, None].astype(np.float64)))):
    fields = dict(common, raw_subints=None, dat_scl=None, dat_offs=None)
    fields.update(extra)
    d = DataBunch(**fields)
    for _ in range(2):
        gt = pptoas.GetTOAs([d], bench.GMODEL, quiet=True); gt.get_TOAs(quiet=True)
    gt = pptoas.GetTOAs([d], bench.GMODEL, quiet=True)
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter(); gt.get_TOAs(quiet=True); dt = time.perf_counter() - t0
    pr.disable()
    print("==", name, "get_TOAs %.1f ms (%.0f TOA/s), fit %.1f ms" % (1e3 * dt, nsub / dt, 1e3 * gt.fit_durations[0]))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
PY
