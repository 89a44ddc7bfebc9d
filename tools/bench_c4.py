"""Config 4: 1 000 000 subints of 512 x 2048 streamed through one GPU in batches
of 10 000 whose data are generated on the device (4 TB of portraits do not fit anywhere);
reports the sustained fit throughput (CUDA-event time of the fit calls only; the synthetic
generator is timed separately) and the recovered-DM statistics over the whole run."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan

NCHAN, NBIN, NU0, BW = 512, 2048, 1500.0, 800.0
P = 1.0 / 345.67890123456789
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
freqs = np.linspace(NU0 - BW / 2 + BW / (2.0 * NCHAN), NU0 + BW / 2 - BW / (2.0 * NCHAN), NCHAN)
gm = os.path.join(ROOT, "tests", "golden", "example.gmodel")
_, _, model = pplib.read_model(gm, pplib.get_bin_centers(NBIN), freqs, P, quiet=True)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(4)
mFT = torch.fft.rfft(torch.from_numpy(model).to(dev), dim=-1)
k = torch.arange(mFT.shape[-1], device=dev, dtype=torch.float64)
nu2 = torch.from_numpy(freqs ** -2.0 - NU0 ** -2.0).to(dev)
data = torch.empty((batch, NCHAN, NBIN), dtype=torch.float32, device=dev)
pl = WidebandPlan(NCHAN, NBIN)
pl.set_model(model.astype(np.float32), freqs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fit_ms = gen_s = 0.0
nconv = 0
pulls, passes, per_batch = [], [], []
done = 0
t_wall = time.perf_counter()
while done < total:
    n = min(batch, total - done)
    t0 = time.perf_counter()
    phi = torch.rand(n, generator=g, device=dev, dtype=torch.float64) - 0.5
    dDM = 3e-4 + 2e-4 * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    for a in range(0, n, 100):
        b = min(n, a + 100)
        sh = -phi[a:b, None] - (pplib.Dconst * dDM[a:b, None] / P) * nu2[None, :]
        ph = torch.exp(2j * np.pi * (sh[:, :, None] * k[None, None, :]))
        clean = torch.fft.irfft(mFT[None] * ph, n=NBIN, dim=-1)
        data[a:b] = clean.to(torch.float32) + 1.5 * torch.randn(clean.shape, generator=g, device=dev, dtype=torch.float32)
    torch.cuda.synchronize()
    gen_s += time.perf_counter() - t0
    e0.record()
    r = pl.fit_batch(data[:n], P, pinned_results=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fit_ms += ms
    per_batch.append(n / ms * 1e3)
    nconv += int((r["return_code"] == 0).sum())
    pulls.append(((r["params"][:, 1] - dDM.cpu().numpy()) / r["param_errs"][:, 1]).copy())
    passes.append(r["nfeval"].mean())
    done += n
pulls = np.concatenate(pulls)
print(json.dumps({"workload": "config 4: %d subints of 512x2048 streamed in batches of %d, data generated on the device" % (total, batch),
                  "TOAs_per_s_fit_only": round(total / fit_ms * 1e3, 1), "fit_s": round(fit_ms / 1e3, 2), "generate_s": round(gen_s, 1),
                  "wall_s": round(time.perf_counter() - t_wall, 1), "converged": nconv, "mean_passes": round(float(np.mean(passes)), 4),
                  "per_batch_TOAs_per_s_min_median_max": [round(float(np.min(per_batch))), round(float(np.median(per_batch))), round(float(np.max(per_batch)))],
                  "dDM_pull_mean": round(float(pulls.mean()), 4), "dDM_pull_rms": round(float(np.sqrt(np.mean(pulls ** 2))), 4),
                  "dDM_pull_max_abs": round(float(np.abs(pulls).max()), 2)}))
