"""Latency of the reference-style per-subint calls (one portrait per call) through the facade."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import synth
from pulseportraiture_b200 import pplib, pptoaslib

out = {}
for (nchan, nbin) in [(64, 512), (512, 2048)]:
    c = synth.make_case(nchan, nbin, 1500., 800., 1)
    data, model, freqs, P = c["data"], c["model"], c["freqs"], c["P"]
    errs = pplib.get_noise(data, chans=True)
    for name, fn in (("fit_portrait", lambda: pplib.fit_portrait(data, model, [0.1, 0.0], P, freqs, errs=errs)),
                     ("fit_portrait_full", lambda: pptoaslib.fit_portrait_full(data, model, [0.1, 0.0, 0, 0, 0], P, freqs, errs=errs,
                                                                                fit_flags=[1, 1, 0, 0, 0], log10_tau=False)),
                     ("fit_phase_shift", lambda: pplib.fit_phase_shift(data.mean(0), model.mean(0)))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            fn()
        out["%s_%dx%d_ms" % (name, nchan, nbin)] = round((time.perf_counter() - t0) / n * 1e3, 3)
print(json.dumps(out))
