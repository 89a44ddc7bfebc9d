"""Throughput of the arbitrary-nbin (Bluestein) path next to the tuned power-of-two one: phi+DM fits of
512-channel portraits, device-resident input."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan

nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
out = []
NBINS = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2048, 2000, 1024, 1000, 1536]
for nbin in NBINS:
    freqs = np.linspace(bench.NU0 - bench.BW / 2 + bench.BW / 1024., bench.NU0 + bench.BW / 2 - bench.BW / 1024., 512)
    _, _, model = pplib.read_model(bench.GMODEL, pplib.get_bin_centers(nbin), freqs, bench.P_EXAMPLE, quiet=True)
    g = torch.Generator(device=dev); g.manual_seed(nbin)
    m = torch.from_numpy(model).to(dev)
    data = (m[None] + 1.5 * torch.randn((nsub,) + model.shape, generator=g, device=dev, dtype=torch.float64)).to(torch.float32)
    with WidebandPlan(512, nbin) as pl:
        pl.set_model(np.ascontiguousarray(model, dtype=np.float64), freqs)
        for _ in range(2):
            r = pl.fit_batch(data, bench.P_EXAMPLE, pinned_results=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            r = pl.fit_batch(data, bench.P_EXAMPLE, pinned_results=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        nconv = int((np.array(r["return_code"]) == 0).sum())   # (pinned results live as long as the plan)
        pl.enable_timing(True)
        pl.fit_batch(data, bench.P_EXAMPLE, pinned_results=True)
        st = pl.stats()
    out.append({"nbin": nbin, "TOAs_per_s": nsub / dt, "ms_per_batch": 1e3 * dt, "ms_spectra": st["ms_spectra"],
                "ms_pass": st["ms_pass"], "x_keep_frac": st["x_keep_frac"], "converged": nconv})
    del data
print(json.dumps({"workload": "%d subints of 512 channels, phi+DM" % nsub, "cases": out}))
