#!/usr/bin/env python
"""Summarise a LIRA_U8_TRACE csv (clock stamps of CTA 0's first 512 (tile, chunk) units of the byte scan's filter pass)."""
import csv, statistics as st, sys
rows = [{k: int(v) for k, v in r.items()} for r in csv.DictReader(open(sys.argv[1]))]
rows = [r for r in rows if r["mma_issued"] > 0 and r["epi_ready"] > 0][50:450]
n = len(rows)
print(n, "units; cycles per unit:", (rows[-1]["mma_issued"] - rows[0]["mma_issued"]) / (n - 1))
d = lambda a, b: st.mean(r[b] - r[a] for r in rows)
print("MMA warp: accumulator free -> operands ready", d("mma_acc_free", "mma_b_ready"), "; operands ready -> issued + committed", d("mma_b_ready", "mma_issued"))
print("MMA warp: previous unit issued -> this accumulator free", st.mean(rows[i]["mma_acc_free"] - rows[i - 1]["mma_issued"] for i in range(1, n)))
print("issued -> epilogue warp 0 sees the accumulator", d("mma_issued", "epi_ready"), "; -> both halves loaded, accumulator released", d("epi_ready", "epi_loaded"))
r2 = [r for r in rows if r["epi_done"] > 0]
print("released -> unit done", st.mean(r["epi_done"] - r["epi_loaded"] for r in r2))
print("epilogue warp 0: unit done -> next accumulator seen", st.mean(rows[i]["epi_ready"] - rows[i - 1]["epi_done"] for i in range(1, n) if rows[i - 1]["epi_done"] > 0))
