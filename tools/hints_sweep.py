"""A/B of the L2 residency hints of the pipelined 3D3V kernel (hd_advection_set_l2_hints) on the 8^6-cell lattice."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
nc = [int(x) for x in os.environ.get("CELLS", "8,8,8,8,8,8").split(",")]
vel = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
op = api.AdvectionOperation(mf, vel, 0.5)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src); ref = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
op.set_l2_hints(0)
op.apply(ref.data_ptr(), src.data_ptr(), 0.0)
import time
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
masks = [int(x) for x in os.environ.get("MASKS", "0,1,2,4,7,7,4,2,1,0,0,7,0,7").split(",")]
for m in masks:
    op.set_l2_hints(m)
    dst.zero_()
    torch.cuda.synchronize()
    time.sleep(float(os.environ.get("SLEEP", "1.5")))  # let the power/clock governor settle: back-to-back runs drift
    ms = timeit(lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0))
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    same = bool(torch.equal(dst, ref))
    print("hints=%d  %.3f ms  %.1f GDoF/s  bit-identical=%s  sm_clock_after=%d MHz power=%.0f W" % (m, ms, n / ms / 1e6, same, clk, pw), flush=True)
