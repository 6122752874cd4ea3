import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
vel = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
def timeit(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-36s %.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
nc = [8] * 6
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
op = api.AdvectionOperation(mf, vel, 0.5)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
ghost = torch.zeros(16, dtype=torch.float64, device="cuda")
A = lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0, ghosts=ghost.data_ptr())
I = lambda: op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_INTERIOR)
for name, fn, sl in (("all", A, 0), ("all", A, 0), ("interior", I, 0), ("interior", I, 0.5), ("all", A, 0.5), ("interior", I, 0.5), ("all x40", None, 0.5)):
    time.sleep(sl)
    if fn is None:
        timeit(name, A, reps=40)
    else:
        timeit(name, fn)
