"""Where the non-kernel time of a fit batch goes: wall time and CUDA-event kernel sums
for several chunk sizes (config 2 shape, device-resident data)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan

NCHAN, NBIN, NU0, BW = 512, 2048, 1500.0, 800.0
P = 1.0 / 345.67890123456789
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
freqs = np.linspace(NU0 - BW / 2 + BW / (2.0 * NCHAN), NU0 + BW / 2 - BW / (2.0 * NCHAN), NCHAN)
gm = os.path.join(ROOT, "tests", "golden", "example.gmodel")
_, _, model = pplib.read_model(gm, pplib.get_bin_centers(NBIN), freqs, P, quiet=True)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
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
pl.set_model(model.astype(np.float32), freqs)
kw = dict(pinned_results=True)
for chunk in [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["512", "1024", "2048", "4000"])]:
    pl.set_chunk(chunk)
    pl.enable_timing(False)
    for _ in range(2):
        r = pl.fit_batch(data, P, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = pl.fit_batch(data, P, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e3
    pl.enable_timing(True)
    r = pl.fit_batch(data, P, **kw)
    st = pl.stats()
    ks = st["ms_spectra"] + st["ms_pass"] + st["ms_update"] + st["ms_guess"]
    print(json.dumps({"chunk": chunk, "nchunks": -(-nsub // chunk), "wall_ms": round(dt, 3), "ms_total_timed": round(st["ms_total"], 3),
                      "kernels_ms": round(ks, 3), "gap_ms": round(dt - ks, 3), "launches": st["launches"],
                      "spectra": round(st["ms_spectra"], 3), "pass": round(st["ms_pass"], 3), "guess": round(st["ms_guess"], 3),
                      "update": round(st["ms_update"], 3), "pass_launches": st["pass_launches"]}))
