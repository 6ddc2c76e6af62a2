#!/usr/bin/env python
"""Turn the raw captures of tools/capture_profiles.sh (gpurun_out/) into the committed summaries
under profiles/: bench JSON, kernel table, ncu launch-list shares, raw CSV pages of the
--set full reports.  Runs without a GPU (ncu -i only reads the reports)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
for src, dst in ((f"{tag}_bench.json", f"{out_tag}_bench_final.json"), (f"{tag}_kernels.json", f"{out_tag}_kernels_final.json"),
                 (f"{tag}_fused_phases.json", f"{out_tag}_fused_phases_final.json"), (f"{tag}_launches.csv", f"{out_tag}_ncu_launches_bench.csv"),
                 (f"{tag}_clocks_idle.csv", f"{out_tag}_clocks.csv"), (f"{tag}_pytest_gpu.log", f"{out_tag}_pytest_gpu.log")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
if os.path.exists(os.path.join(G, f"{tag}_launches.csv")):
    for skip, name in ((0, "all"), (None, "steady_state")):
        # steady state = the last 15 timestamps' launches (3 warm + 3 timed + 3 + 3 e2e + 3 profiled)
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), os.path.join(G, f"{tag}_launches.csv"), "0"],
                             capture_output=True, text=True).stdout
        if skip is None:
            txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), os.path.join(G, f"{tag}_launches.csv"), "after:k_gen_orders"],
                                 capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out_tag}_ncu_launches_{name}_summary.txt"), "w").write(txt)
for k in ("fused", "probe", "mrg", "onesweep", "probe_bulk"):
    rep = os.path.join(G, f"{tag}_{k}.ncu-rep")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out_tag}_ncu_full_{k}_raw.csv"), "w").write(raw)
print(sorted(os.listdir(P)))
