"""Summarise a Q3TTS_CODEC_UNIT_TRACE dump: per-tile phase times (cycles) of the fused residual-unit kernel, median over CTAs."""
import json, sys
import numpy as np
d = json.load(open(sys.argv[1]))
S = d.get("slots", 8)
s = np.array(d["stamps"], dtype=np.int64).reshape(d["ctas"], d["tiles"], S)
names = ["c7_issue_start", "c7_issued", "c1_issue_start", "c7_done_seen", "e1_done", "c1_done_seen", "e2_done", "P_tile_start", "P_tile_requested", "unused9", "unused10", "unused11"][:S]
t0 = s[:, 0, 0][:, None, None]
rel = s - t0
print(f"C={d['C']} dil={d['dil']} ctas={d['ctas']} b_stages={d.get('b_stages')} a_stages={d.get('a_stages')}")
for ti in range(d["tiles"]):
    row = {n: int(np.median(rel[:, ti, i])) for i, n in enumerate(names) if (s[:, ti, i] > 0).all()}
    print(ti, row)
per = np.median(np.diff(s[:, :, 6], axis=1), axis=0)
print("e2_done period per tile (cycles, median over CTAs):", per.astype(int).tolist())
