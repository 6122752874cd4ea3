"""Where does the pipelined kernel's time go?  Variants: L2 hints, and velocities with some directions switched off
(no face loads for those).  Run plain for timings, under `ncu --metrics dram__bytes_read.sum,...` for traffic."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
nc = [8] * 6
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
reps = int(os.environ.get("REPS", "10"))
V = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
cases = [("full hints0", V, 0), ("full hints1", V, 1), ("full hints7", V, 7), ("a5=0", V[:5] + (0.0,), 0), ("a4=a5=0", V[:4] + (0.0, 0.0), 0),
         ("a3=a4=a5=0", V[:3] + (0.0,) * 3, 0), ("a1..a5=0", (1.0, 0, 0, 0, 0, 0), 0), ("a=0", (0.0,) * 6, 0)]
for name, vel, hints in cases:
    op = api.AdvectionOperation(mf, vel, 0.5)
    op.set_l2_hints(hints)
    for _ in range(2 if reps > 1 else 1):
        op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
    torch.cuda.synchronize()
    if reps > 1:
        time.sleep(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%-14s %.3f ms  %.1f GDoF/s" % (name, ms, n / ms / 1e6), flush=True)
