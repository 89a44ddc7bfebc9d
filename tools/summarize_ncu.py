"""Turn ncu artefacts brought back in gpurun_out/ into the tracked summaries
under profiles/ (launch list shares, per-kernel raw metrics, top stalled SASS).

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/summarize_ncu.py kernel gpurun_out/prof_pass_b.ncu-rep profiles/r01_k_pass2.md
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    for r in rows[1:]:
        try:
            name, val, u = r[ix["Kernel Name"]], float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        except (ValueError, IndexError):
            continue
        if u not in unit:
            continue
        key = name.replace("void ", "").split("(")[0]
        a = agg.setdefault(key, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += val * unit[u]
    ours = {k: v for k, v in agg.items() if k.startswith("ppb::") or k.startswith("k_")}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        fh.write("Source: `%s`.  Times are cold-cache and serialised by the profiler: "
                 "compare SHARES.  Only this repo's kernels (namespace ppb) are listed; the "
                 "torch kernels in the same capture generate the synthetic input (untimed setup).\n\n" % src)
        fh.write("| kernel | launches | grid (last) | block | total ms | share |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %s | %s | %.3f | %.1f%% |\n" % (k, v[0], v[2], v[3], v[1], 100 * v[1] / tot))
        fh.write("\nTotal (our kernels): %.3f ms in %d launches.\n" % (tot, sum(v[0] for v in ours.values())))


def kernel(rep, dst):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full summary: `%s`\n" % rep]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append("\n## %s (launch id %s)\n\n| metric | value | unit |\n|---|---|---|" % (d.get("Kernel Name", "?"), d.get("ID")))
        for m in RAW:
            if m in d:
                out.append("| %s | %s | %s |" % (m, d[m], u.get(m, "")))
        out.append("\nWarps stalled per issue-active cycle (> 0.3):\n")
        for k, v in d.items():
            if "average_warps_issue_stalled" in k and "per_issue_active" in k:
                try:
                    if float(v) > 0.3:
                        out.append("- %s = %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
                except ValueError:
                    pass
    src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["# Samples"] and r[ix["# Samples"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    out.append("\n## Top stalled SASS instructions (warp-stall samples, first launch)\n\n| samples | share | SASS |\n|---|---|---|")
    seen = 0
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:18]:
        n = int(r[ix["# Samples"]])
        seen += n
        out.append("| %d | %.1f%% | `%s` |" % (n, 100.0 * n / tot, r[ix["Source"]].strip()))
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
