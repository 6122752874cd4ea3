"""Timeline of one CTA of the three-round kernel (HD_R6_TRACE build, tools/build_variants.sh r6_trace "-DHD_R6_TRACE"):
runs one apply on the bench lattice with HD_R6_TRACE_FILE set and prints, per cell, when each round/warp passed its
barriers (cycles relative to the CTA's first event) and how long it waited at each.
usage: HD_LIBHDGPU=hyperdeal_b200/lib/variants/libhdgpu_r6_trace.so python tools/r6_timeline.py [out.txt]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tf = "/tmp/r6_trace.bin"
os.environ["HD_R6_TRACE_FILE"] = tf
import torch
from hyperdeal_b200 import api
ctx = api.Context(0)
mf = api.MatrixFree(ctx, 3, 3, 3, [8] * 6, (0.0,) * 6, (1.0,) * 6)
src = torch.randn(mf.n_dofs, dtype=torch.float64, device="cuda"); dst = torch.zeros_like(src)
vel = tuple(float(x) for x in os.environ.get("AB_VEL", "1.0,0.15,-0.05,0.1,-0.15,0.5").split(","))
op = api.AdvectionOperation(mf, vel, 0.5)
for _ in range(3):
    op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
torch.cuda.synchronize()
t = np.fromfile(tf, dtype=np.int64).reshape(13, -1, 16)
ncell = t.shape[1]
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
names = ["r0", "r1", "r2"]
# per round (warp 0 of each): cell start, waits
print("cell | producer: wait_empty issue | r0: top wait_pEmpty wait_fullU end | r1: top wait_fullU wait_p0(t0) end | r2: top wait_fullU wait_p1 end   (cycles; waits = durations)", file=out)
for k in range(40, 140):
    row = ["%4d" % k]
    p = rel[12, k]
    row.append("P %8d %6d %6d" % (p[0], p[1] - p[0], p[2] - p[1]))
    for R in range(3):
        e = rel[4 * R, k]
        row.append("%s %8d pw %5d fu %5d pp %5d end %8d" % (names[R], e[0], e[1] - e[0], e[3] - e[2], (e[5] - e[4]) if e[5] > 0 else 0, e[7]))
    print(" | ".join(row), file=out)
# summary over cells 50..450
for R in range(3):
    for w in range(4):
        e = rel[4 * R + w, 50:450]
        dur = np.diff(e[:, 0]).mean()
        print("%s warp %d: cell period %.0f cyc, pEmpty wait %.0f, fullU(k+1) wait %.0f, partial wait %.0f" % (names[R], w, dur, (e[:, 1] - e[:, 0]).mean(), (e[:, 3] - e[:, 2]).mean(), (e[:, 5] - e[:, 4]).clip(min=0).mean()), file=out)
e = rel[12, 50:450]
print("producer: period %.0f, emptyU wait %.0f, issue work %.0f; lead of TMA issue over r0's start of the same cell: %.0f cycles" % (np.diff(e[:, 0]).mean(), (e[:, 1] - e[:, 0]).mean(), (e[:, 2] - e[:, 1]).mean(), (rel[0, 50:450, 0] - e[:, 2]).mean()), file=out)
for R in range(3):
    print("%s start of cell k minus TMA issue of cell k: %.0f cycles; minus r2 end of cell k-1: %.0f" % (names[R], (rel[4 * R, 50:450, 0] - rel[12, 50:450, 2]).mean(), (rel[4 * R, 50:450, 0] - rel[8, 49:449, 7]).mean()), file=out)

# phases inside the two tasks of a cell (events 8..15): trace terms | request | main terms | epilogue, per round, warp 0 and 1
print("phase durations (cycles, mean over cells 50..450): [task j] top->traces | traces->request issued | request->main done | main->task end", file=out)
for R in range(3):
    for w in (0, 1):
        e = rel[4 * R + w, 50:450].astype(float)
        a0 = e[:, 8] - e[:, 1]
        b0 = e[:, 9] - e[:, 8]
        c0 = e[:, 10] - e[:, 9]
        d0 = e[:, 11] - e[:, 10]
        a1 = e[:, 12] - e[:, 11]
        b1 = e[:, 13] - e[:, 12]
        c1 = e[:, 14] - e[:, 13]
        d1 = e[:, 15] - e[:, 14]
        print("%s warp %d: task0 %5.0f %5.0f %5.0f %5.0f | task1 %5.0f %5.0f %5.0f %5.0f | cell %5.0f" % (names[R], w, a0.mean(), b0.mean(), c0.mean(), d0.mean(), a1.mean(), b1.mean(), c1.mean(), d1.mean(), (e[:, 15] - e[:, 0]).mean()), file=out)
