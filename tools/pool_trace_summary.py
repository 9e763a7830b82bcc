"""Summarise gpurun_out/pool_trace_<tag>.npy (tools/pool_probe.py): per-CTA tile period and mean stage hand-over latencies."""
import sys
import numpy as np
tr = np.load("gpurun_out/pool_trace_%s.npy" % sys.argv[1])
names = "issue,land,mma_issue,ready,numerators,weights,pooled"
for c in range(tr.shape[0]):
    x = tr[c]; n = int((x[:, 7] > 0).sum())
    if n < 8:
        continue
    d = np.diff(x[2:n, 7])
    print("  CTA %d: %d tiles, first tile pooled after %d cycles, period mean %.0f; stage deltas (%s): %s"
          % (c, n, x[0, 7] - x[0, 0], d.mean(), names, np.round((x[4:n, 1:] - x[4:n, :-1]).mean(0)).astype(int).tolist()))
