"""Row-tile sweep of the pipelined 3D3V kernel on the 8^6-cell bench lattice: time per apply for a set of tile shapes
(hd_advection_set_row_tile), checks that every order gives bit-identical results.  Run plain for timings; under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (REPS=1) for the traffic of each shape."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
nc = [int(os.environ.get("CELLS", "8"))] * 6
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src); ref = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
reps = int(os.environ.get("REPS", "10"))
V = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
shapes = [(0, 0, 0, 0, 0), (4, 4, 4, 4, 0), (4, 4, 4, 2, 0), (4, 4, 2, 2, 0), (2, 4, 4, 4, 0), (8, 4, 4, 2, 0), (8, 4, 4, 4, 0), (2, 2, 4, 4, 0), (2, 2, 2, 2, 0), (4, 4, 4, 4, 4),
          (4, 2, 2, 4, 0), (8, 8, 2, 2, 0), (4, 4, 4, 8, 0), (8, 2, 2, 2, 0)]
if os.environ.get("SHAPES"):
    shapes = [tuple(int(x) for x in s.split(",")) for s in os.environ["SHAPES"].split(";")]
op = api.AdvectionOperation(mf, V, 0.5)
op.set_row_tile((0, 0, 0, 0, 0))
op.apply(ref.data_ptr(), src.data_ptr(), 0.0)
torch.cuda.synchronize()
for shp in shapes:
    op.set_row_tile(shp)
    for _ in range(2 if reps > 1 else 1):
        op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
    torch.cuda.synchronize()
    same = bool(torch.equal(dst, ref))
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        import pynvml
        pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
        print("tile %-16s %.3f ms  %.1f GDoF/s  identical=%s  sm_clock_after=%d MHz power=%.0f W" % (shp, ms, n / ms / 1e6, same, pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3), flush=True)
    else:
        print("tile %-16s identical=%s" % (shp, same), flush=True)
