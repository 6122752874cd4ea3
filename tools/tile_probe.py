"""2D2V k=3 tile kernel on the 64x64x32x32 lattice (BASELINE.json configs[0]): ms per apply.  HD_TILE_THREADS selects the CTA size."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
dt = np.float32 if os.environ.get("F32") else np.float64
mf = api.MatrixFree(ctx, 2, 2, 3, [64, 64, 32, 32], (0.0,) * 4, (1.0,) * 4, dtype=dt)
op = api.AdvectionOperation(mf, (1.0, 0.15, -0.05, 0.1), 0.5)
src = torch.empty(mf.n_dofs, dtype=torch.float32 if os.environ.get("F32") else torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
reps = int(os.environ.get("REPS", "10"))
for _ in range(2): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("HD_TILE_THREADS=%s kernel=%s %.3f ms %.1f GDoF/s" % (os.environ.get("HD_TILE_THREADS", "128") + (" f32" if os.environ.get("F32") else ""), op.kernel_name, ms, mf.n_dofs / ms / 1e6), flush=True)
