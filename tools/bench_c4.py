"""Config 4 (SURVEY 8d / BASELINE.json configs[3]): 1 000 000 subints of 512 x 2048, sharded across the
ranks of one node (torchrun, one rank per GPU: contiguous ranges of the global campaign, multigpu.shard_range)
and streamed through each GPU in batches of 10 000 whose data are generated on the device (4 TB of portraits
fit nowhere).  After every batch the TOA-level results of all ranks are gathered on rank 0 through shared
memory (multigpu.SharedGather).  Reports the sustained fit throughput of the whole job (device time of the fit
calls, max over ranks; the synthetic generator is timed separately) and the recovered-DM statistics of the
gathered campaign.

    python tools/bench_c4.py [total] [batch]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4.py [total] [batch]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from pulseportraiture_b200 import pplib
from pulseportraiture_b200.engine import WidebandPlan
from pulseportraiture_b200.multigpu import shard_range, SharedGather, bind_to_gpu_numa

NCHAN, NBIN, NU0, BW = 512, 2048, 1500.0, 800.0
P = 1.0 / 345.67890123456789
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
gloo = None
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")
torch.cuda.set_device(local)
bind_to_gpu_numa(local)
dev = torch.device("cuda", local)
freqs = np.linspace(NU0 - BW / 2 + BW / (2.0 * NCHAN), NU0 + BW / 2 - BW / (2.0 * NCHAN), NCHAN)
gm = os.path.join(ROOT, "tests", "golden", "example.gmodel")
_, _, model = pplib.read_model(gm, pplib.get_bin_centers(NBIN), freqs, P, quiet=True)
g = torch.Generator(device=dev); g.manual_seed(4 + rank)
mFT = torch.fft.rfft(torch.from_numpy(model).to(dev), dim=-1)
k = torch.arange(mFT.shape[-1], device=dev, dtype=torch.float64)
nu2 = torch.from_numpy(freqs ** -2.0 - NU0 ** -2.0).to(dev)
data = torch.empty((batch, NCHAN, NBIN), dtype=torch.float32, device=dev)
pl = WidebandPlan(NCHAN, NBIN, device=local)
pl.set_model(np.ascontiguousarray(model, dtype=np.float64), freqs)
a0, b0 = shard_range(total, rank, world)                       # this rank's part of the campaign
shg = SharedGather(batch, 6, group=gloo, tag="c4") if world > 1 else None
# untimed warm-up (buffer allocation, first launches) on a noise-only batch of the campaign's shape
data.normal_(0.0, 1.5, generator=g)
for _ in range(2):
    pl.fit_batch(data, P, pinned_results=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fit_ms = gen_s = gather_s = 0.0
rows, per_batch = [], []
done = a0
t_wall = time.perf_counter()
nbatches = -(-(shard_range(total, 0, world)[1]) // batch)     # rank 0 has the longest shard: everyone loops as often
for ib in range(nbatches):
    n = max(0, min(batch, b0 - done))
    t0 = time.perf_counter()
    phi = torch.rand(max(n, 1), generator=g, device=dev, dtype=torch.float64) - 0.5
    dDM = 3e-4 + 2e-4 * torch.randn(max(n, 1), generator=g, device=dev, dtype=torch.float64)
    for a in range(0, n, 100):
        b = min(n, a + 100)
        sh = -phi[a:b, None] - (pplib.Dconst * dDM[a:b, None] / P) * nu2[None, :]
        ph = torch.exp(2j * np.pi * (sh[:, :, None] * k[None, None, :]))
        clean = torch.fft.irfft(mFT[None] * ph, n=NBIN, dim=-1)
        data[a:b] = clean.to(torch.float32) + 1.5 * torch.randn(clean.shape, generator=g, device=dev, dtype=torch.float32)
    torch.cuda.synchronize()
    gen_s += time.perf_counter() - t0
    pack = np.zeros((0, 6))
    if n:
        e0.record()
        r = pl.fit_batch(data[:n], P, pinned_results=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fit_ms += ms
        per_batch.append(n / ms * 1e3)
        pack = np.stack([r["params"][:, 0], r["param_errs"][:, 0], r["params"][:, 1], r["param_errs"][:, 1],
                         (r["params"][:, 1] - dDM[:n].cpu().numpy()) / r["param_errs"][:, 1],
                         r["nfeval"] + 1000.0 * (r["return_code"] != 0)], axis=1)
    t0 = time.perf_counter()
    allrows = shg.gather(pack) if world > 1 else pack            # host-side gather of this batch's TOA-level rows
    gather_s += time.perf_counter() - t0
    if rank == 0:
        rows.append(allrows)
    done += n
wall = time.perf_counter() - t_wall
tm = torch.tensor([fit_ms, gather_s], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    rows = np.concatenate(rows)
    pulls, nfe = rows[:, 4], rows[:, 5]
    fit_s, gat_s = float(tm[0]) / 1e3, float(tm[1])
    print(json.dumps({"workload": "config 4: %d subints of 512x2048 sharded over %d GPU(s), streamed in batches of %d generated on the device, "
                                  "TOA-level results gathered on rank 0 through shared memory after every batch" % (total, world, batch),
                      "n_gpus": world, "TOAs_per_s_fit_only": round(total / fit_s, 1), "TOAs_per_s_fit_plus_gather": round(total / (fit_s + gat_s), 1),
                      "fit_s_max_over_ranks": round(fit_s, 3), "gather_s": round(gat_s, 3), "generate_s": round(gen_s, 1), "wall_s": round(wall, 1),
                      "gathered_rows": int(len(rows)), "converged": int((nfe < 1000).sum()), "mean_passes": round(float(np.mean(nfe % 1000)), 4),
                      "rank0_per_batch_TOAs_per_s_min_median_max": [round(float(np.min(per_batch))), round(float(np.median(per_batch))), round(float(np.max(per_batch)))],
                      "dDM_pull_mean": round(float(pulls.mean()), 4), "dDM_pull_rms": round(float(np.sqrt(np.mean(pulls ** 2))), 4),
                      "dDM_pull_max_abs": round(float(np.abs(pulls).max()), 2)}))
if world > 1:
    shg.close()
    dist.destroy_process_group()
