#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r02_facade_probe.txt
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from pulseportraiture_b200.engine import WidebandPlan
freqs, model = bench.make_model()
dev = torch.device('cuda', 0)
phi, dDM = bench.global_draws(512)
data = bench.make_device_batch(model, freqs, phi, dDM, 1, dev)
print(json.dumps(bench.facade_timing(data, freqs, model, 256)))
# pageable float32 / float64 / pinned through the C ABI
pl = WidebandPlan(bench.NCHAN, bench.NBIN); pl.set_model(model.astype(np.float32), freqs)
h32 = data.cpu().numpy(); h64 = h32.astype(np.float64)
for name, arr in (("pageable f32", h32), ("pageable f64", h64)):
    pl.fit_batch(arr, bench.P_EXAMPLE)
    t = time.perf_counter(); r = pl.fit_batch(arr, bench.P_EXAMPLE); dt = time.perf_counter() - t
    print(name, "%.0f TOA/s  %.1f GB/s" % (len(arr) / dt, arr.nbytes / dt / 1e9), int((r["return_code"] == 0).sum()))
PY
