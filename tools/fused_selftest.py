"""1-GPU timing of the fused-halo kernel with a self exchange (brick whose periodic neighbour is itself) at full size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
ctx = api.Context(0)
dirs = tuple(int(x) for x in os.environ.get("DIRS", "2").split(",") if x != "")
nc = [int(x) for x in os.environ.get("CELLS", "8,8,8,8,8,8").split(",")]
vel = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)
sk = [[api.SIDE_GHOST] * 2 if d in dirs else [api.SIDE_PERIODIC_LOCAL] * 2 for d in range(6)]
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6, side_kind=sk)
op = api.AdvectionOperation(mf, vel, 0.5)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
ghost = torch.zeros(mf.halo_total, dtype=torch.float64, device="cuda")
counters = torch.zeros(12, dtype=torch.int32, device="cuda")
needed = op.ghost_sides()
sends = [(d, s, ghost.data_ptr() + 8 * mf.halo_offset(d, 1 - s), counters.data_ptr() + 4 * (2 * d + (1 - s))) for d in dirs for s in range(2) if needed[2 * d + (1 - s)]]
def timeit(name, fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-36s %.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
arrived = [0]  # what every arrival counter holds after the applications so far
def fused():
    arrived[0] += op.n_halo_senders
    op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), sends, counters.data_ptr(), arrived[0])
mask = [0] * 12
for (d, s, _, _) in sends: mask[2 * d + s] = 1
send = torch.zeros(mf.halo_total, dtype=torch.float64, device="cuda")
print("dirs", dirs, "cells", nc, "halo MB", sum(mf.ghost_size(d, s) for (d, s, _, _) in sends) * 8 / 1e6)
timeit("apply all (ghosts in place)", lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0, ghosts=ghost.data_ptr()))
timeit("interior", lambda: op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_INTERIOR))
timeit("boundary", lambda: op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_BOUNDARY))
timeit("pack kernel", lambda: mf.halo_pack(src.data_ptr(), send.data_ptr(), send_mask=mask))
for ns in [int(x) for x in os.environ.get("SENDERS", "16,32,64,148").split(",")]:
    op.set_halo_senders(ns)
    timeit("fused (pack+interior+wait+boundary), %3d sender CTAs" % op.n_halo_senders, fused)
# structure only: no sends, arrival counters already at their target -> interior list, then boundary list in one launch
counters.fill_(1 << 30)
timeit("fused structure only (no sends, flags preset)", lambda: op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), [], counters.data_ptr(), 1))
assert not op.overlap_timed_out()
