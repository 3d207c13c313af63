#!/usr/bin/env python
"""Per-warp durations of nn_grid_coop_kernel at the pass the instrumented build names (-DICPB_COOP_CLOCKS=<pass>, see
tools/README.md)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))
import numpy as np, icpb200
from icpb200 import synth
d0, d1, col, _ = synth.frame_pair()
ctx = icpb200.Context(0)
K = icpb200.reference_intrinsics_v1(); cam = np.array([5, 5, 5], np.float32)
h, w = d0.shape
tgt, dat = ctx.cloud(w * h), ctx.cloud(w * h)
tgt.from_depth(d0, col, K); tgt.transform(None, cam)
dat.from_depth(d1, col, K); dat.transform(None, cam)
res, _, _ = ctx.icp_register(dat, tgt, 6, 0.0, 0.75, 0, nn_mode=1)
nw = (dat.n + 31) // 32
buf = np.zeros(16384 * 8, np.int32)
rc = ctx.lib.icpb_debug_coop_clocks(buf.ctypes.data_as(C.c_void_p), buf.size)
assert rc == 0
t = buf.reshape(-1, 8)[:nw]
cyc = t[:, 0].astype(np.int64)
print("warps", nw, "cycles: mean %.0f median %.0f p90 %.0f p99 %.0f max %d  sum/1e6 %.1f" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max(), cyc.sum() / 1e6))
start = (t[:, 6].astype(np.int64) - t[:, 6].min()) & 0x7fffffff          # globaltimer, ns
dur = (t[:, 7].astype(np.int64) >> 8) & 0xfffff
sm = t[:, 7] & 0xff
end = start + dur
span = end.max()
print("kernel span %d ns, last warp start %d ns, mean warp %d ns, max warp %d ns" % (span, start.max(), dur.mean(), dur.max()))
order = np.argsort(-cyc)[:12]
for o in order:
    print("warp %5d cycles %7d ns %6d rows %5d cells %5d box %6d staged %5d start %7d sm %d" % (o, cyc[o], dur[o], t[o, 2], t[o, 3], t[o, 4], t[o, 5], start[o], sm[o]))
for name, col_ in (("staged", 5), ("cells", 3), ("rows", 2)):
    print("corr(cycles, %s) = %.3f" % (name, np.corrcoef(cyc, t[:, col_])[0, 1]))
# concurrency over time
ts = np.linspace(0, span, 21)
for a, b in zip(ts[:-1], ts[1:]):
    mid = 0.5 * (a + b)
    act = ((start <= mid) & (end > mid)).sum()
    print("t=%6d ns active warps %5d (%.1f per SM)" % (mid, act, act / 148.0))
last_end_sm = np.array([end[sm == s_].max() if (sm == s_).any() else 0 for s_ in range(148)])
print("per-SM last end: min %d median %d max %d" % (last_end_sm.min(), np.median(last_end_sm), last_end_sm.max()))
per_sm_n = np.bincount(sm, minlength=148)
print("warps per SM: min %d max %d" % (per_sm_n.min(), per_sm_n.max()))
