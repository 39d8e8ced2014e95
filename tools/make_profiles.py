#!/usr/bin/env python
"""Turn the raw outputs of tools/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
   tools/make_profiles.py <tag> <round-prefix>     e.g.  tools/make_profiles.py r1b r1"""
import csv, hashlib, json, os, subprocess, sys
tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- bench line -----------------------------------------------------------------------------
line = open(os.path.join(G, f"{tag}_bench.json")).read().strip().splitlines()[-1]
open(os.path.join(P, f"{rnd}_bench_line.json"), "w").write(open(os.path.join(G, f"{tag}_bench.err")).read() + line + "\n")

# ---- launch list + one-step breakdown ----------------------------------------------------------
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
with open(os.path.join(P, f"{rnd}_launches.csv"), "w") as f:
    csv.writer(f).writerows(rows[hi:])
h, data = rows[hi], rows[hi + 1:]
kn, mv, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
L = [(r[kn].split("(")[0].replace("void ", "").replace("lira::", ""), float(r[mv].replace(",", "")), r[gs]) for r in data if len(r) > mv]
idx = [i for i, (n, _, _) in enumerate(L) if n.startswith("prep_queries")]
step = L[idx[-2]:]
step = step[:[i for i, (n, _, _) in enumerate(step) if n.startswith("refine_topk")][0] + 1]
tot = sum(t for _, t, _ in step)
with open(os.path.join(P, f"{rnd}_step_breakdown.txt"), "w") as f:
    f.write(f"# one step of bench.py (config 1, 10k queries) from profiles/{rnd}_launches.csv: ncu gpu__time_duration.sum per launch\n"
            "# (cold-cache, serialised: compare shares)\n")
    for n, t, g in step:
        f.write(f"{n[:48]:48s} {t / 1000:8.1f} us {100 * t / tot:5.1f} %  grid {g}\n")
    f.write(f"{'total':48s} {tot / 1000:8.1f} us\n")

# ---- full capture of the scan kernels --------------------------------------------------------------
rep = os.path.join(G, f"{tag}_tc_scan.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units = r[0], r[1]
pats = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct"]
traffic = None
with open(os.path.join(P, f"{rnd}_scan_ncu_metrics.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none, bench.py config 1 (gpurun_out/{tag}_tc_scan.ncu-rep): seed pass <1,0> and filter pass <0,0>\n")
    for row in r[2:]:
        d = dict(zip(hdr, row))
        f.write("== " + d["Kernel Name"][:80] + "\n")
        for k in hdr:
            if any(p in k for p in pats) and d[k] not in ("", "n/a"):
                f.write(f"   {k} = {d[k]} {units[hdr.index(k)]}\n")
        if "<0, 0>" in d["Kernel Name"] or "(bool)0, (bool)0" in d["Kernel Name"]:
            def b(name):
                v, u = float(d[name].replace(",", "")), units[hdr.index(name)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
            hsh = hashlib.sha256()
            cs = os.path.join(ROOT, "lira-ann-search_b200", "csrc")
            for fn in sorted(os.listdir(cs)):
                if fn.endswith(".cuh"):   # (the kernels live in the .cuh files; lira_b200.cu is host code)
                    hsh.update(open(os.path.join(cs, fn), "rb").read())
            traffic = {"kernel": "tc_scan_kernel<false, false>", "workload": "sift1m-shape", "source_hash": hsh.hexdigest()[:16],
                       "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "algorithmic_bytes": json.loads(line)["roofline"]["algorithmic_bytes"],
                       "source": f"ncu --set full --clock-control none (gpurun_out/{tag}_tc_scan.ncu-rep): dram__bytes_read.sum + dram__bytes_write.sum of one "
                                 "launch, bench.py config 1, 10k-query batch. The kernel streams the fp16 shadow copy of the list rows (half the "
                                 "fp32 algorithmic bytes) plus the 32-byte augmented-norm block per row; lists probed by more than 128 queries "
                                 "are read once per query tile, the repeats mostly from L2."}
if traffic:
    json.dump(traffic, open(os.path.join(P, f"{rnd}_scan_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{rnd}_step_breakdown.txt")).read())
print(json.dumps(traffic, indent=1))
