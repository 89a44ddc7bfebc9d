"""Compact table of tools/bench_c3.py JSON lines read from stdin."""
import sys, json
for l in sys.stdin:
    if not l.startswith("{"):
        print(l.rstrip()); continue
    d = json.loads(l)
    print(d["workload"][22:37], "frac", d["coarse_frac"], "ms", round(d["ms_per_batch"], 2), "TOA/s", round(d["TOAs_per_s"]),
          "eval", round(d["mean_evaluations"], 2), "full", d["pass_launches"], "coarse", d["coarse_launches"],
          "ms_c", round(d["ms_coarse"], 2), "ms_p", round(d["ms_pass"], 2), "ms_u", round(d["ms_update"], 2),
          "ms_s", round(d["ms_spectra"], 2), "dpar", d.get("max_dparam_sigma_vs_first"), "pull", round(d["dDM_pull_rms"], 3))
