#!/usr/bin/env python
"""Summarise a LIRA_TC_TRACE csv (per-chunk SM clock stamps of CTA 0 of the tensor-core scan)."""
import csv, statistics as st, sys
rows = [{k: int(v) for k, v in r.items()} for r in csv.DictReader(open(sys.argv[1]))]
rows = [r for r in rows if r["mma_issued"] > 0]
n = len(rows)
t0 = rows[0]["prod_start"]
print(n, "chunks traced")
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (30, 60)
for r in rows[lo:hi]:
    f = lambda k: r[k] - t0
    print(f"{r['chunk']:4d} prod {f('prod_start'):8d} accfree {f('mma_acc_free'):8d} b0 {f('mma_b0'):8d} blast {f('mma_blast'):8d} issued {f('mma_issued'):8d} |"
          f" e4 {f('epi4_ready'):8d}-{f('epi4_done'):8d} e8 {f('epi8_ready'):8d}-{f('epi8_done'):8d} | epi4 {r['epi4_done']-r['epi4_ready']:5d} epi8 {r['epi8_done']-r['epi8_ready']:5d} rdy-iss {r['epi4_ready']-r['mma_issued']:5d}")
print("cycles per chunk", (rows[-1]["mma_issued"] - rows[0]["mma_issued"]) / (n - 1))
print("epi busy", st.mean(r["epi4_done"] - r["epi4_ready"] for r in rows), st.mean(r["epi8_done"] - r["epi8_ready"] for r in rows))
print("mma: wait acc", st.mean(rows[i]["mma_acc_free"] - rows[i - 1]["mma_issued"] for i in range(1, n)),
      "accfree->b0", st.mean(r["mma_b0"] - r["mma_acc_free"] for r in rows), "b0->blast", st.mean(r["mma_blast"] - r["mma_b0"] for r in rows),
      "blast->issued", st.mean(r["mma_issued"] - r["mma_blast"] for r in rows))
print("issued -> epi ready", st.mean(r["epi4_ready"] - r["mma_issued"] for r in rows))
print("producer period", (rows[-1]["prod_start"] - rows[0]["prod_start"]) / (n - 1))
