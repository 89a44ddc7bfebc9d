"""Do two plans on two streams of ONE device overlap usefully (k_spectra of one batch with
k_pass2 of the other)?  Aggregate TOAs/s of 1 plan x 2n subints vs 2 threads x n subints."""
import os, sys, time, json, threading
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

_streams = []
def make_plan(chunk):
    pl = WidebandPlan(NCHAN, NBIN)
    _streams.append(torch.cuda.Stream(device=dev))
    pl.set_stream(_streams[-1])
    pl.set_model(model.astype(np.float32), freqs)
    pl.set_chunk(chunk)
    return pl

def timed(fn, reps=4):
    fn(); fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

for chunk in (1000, 500):
    p0, p1 = make_plan(chunk), make_plan(chunk)
    half = nsub // 2
    def single():
        p0.fit_batch(data, P, pinned_results=True)
    def dual():
        th = [threading.Thread(target=lambda pl=pl, d=d: pl.fit_batch(d, P, pinned_results=True))
              for pl, d in ((p0, data[:half]), (p1, data[half:]))]
        for t in th: t.start()
        for t in th: t.join()
    t1 = timed(single); t2 = timed(dual)
    print(json.dumps({"chunk": chunk, "single_plan_TOAs_per_s": round(nsub / t1), "two_plans_TOAs_per_s": round(nsub / t2),
                      "single_ms": round(t1 * 1e3, 2), "dual_ms": round(t2 * 1e3, 2)}))
    del p0, p1
