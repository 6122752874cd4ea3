"""Times the pieces of one multi-GPU step in isolation (run under torchrun; rank 0 prints)."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
from hyperdeal_b200.partition import BrickPartition, HaloExchange, PeerHaloExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
order = tuple(int(x) for x in os.environ.get("SPLIT", "2,1,0").split(","))
part = BrickPartition(world, rank, [8] * 6, split_order=order)
if os.environ.get("LAYOUT", "x") == "xv":
    recipe = BrickPartition(world, rank, [8] * 6, split_order=(2, 1, 0))
    cut = BrickPartition(world, rank, [8] * 6, split_order=(2, 1, 5))
    part = BrickPartition(world, rank, [g // c for g, c in zip(recipe.n_cells_global, cut.grid)], grid=cut.grid)
ctx = api.Context(local)
mf = api.MatrixFree(ctx, 3, 3, 3, part.n_cells, (0.0,) * 6, (1.0,) * 6, n_cells_global=part.n_cells_global, cell_offset=part.cell_offset, side_kind=part.side_kind)
op = api.AdvectionOperation(mf, (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
halo = mf.halo_total
send = torch.empty(halo, dtype=torch.float64, device="cuda"); ghost = torch.zeros(halo, dtype=torch.float64, device="cuda")
offsets = {(d, s): mf.halo_offset(d, s) for d in range(6) for s in range(2)}
sizes = {(d, s): mf.ghost_size(d, s) for d in range(6) for s in range(2)}
ex = HaloExchange(part, offsets, sizes, op.ghost_sides())
peer = PeerHaloExchange(part, offsets, sizes, halo, op.ghost_sides(), torch.device("cuda", local), ctx=ctx)
mask = ex.send_mask()

def timeit(name, fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%-28s %.3f ms" % (name, t.item()), flush=True)

if rank == 0: print("grid", part.grid, "halo MB sent", ex.bytes_per_exchange[0] * 8 / 1e6)
timeit("pack local", lambda: mf.halo_pack(src.data_ptr(), send.data_ptr(), send_mask=mask))
timeit("nccl exchange", lambda: HaloExchange.finish(ex.start(send, ghost)))
timeit("pack peer (no barrier)", lambda: mf.halo_pack(src.data_ptr(), None, send_mask=peer.mask, peer_dst=peer.peer_dst[0]))
def pack_signal():
    g, e = peer.start(mf, ctx, src.data_ptr())
    peer.wait_ready(ctx, e)
    peer.consumed(ctx)
timeit("pack peer + flags", pack_signal)
timeit("apply all", lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0, ghosts=ghost.data_ptr()))
timeit("apply interior", lambda: op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_INTERIOR))
timeit("apply boundary", lambda: op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, ghost.data_ptr(), api.PART_BOUNDARY))
def serial():
    g, e = peer.start(mf, ctx, src.data_ptr())
    op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_INTERIOR)
    peer.wait_ready(ctx, e)
    op.apply_part(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), api.PART_BOUNDARY)
    peer.consumed(ctx)
def fused():
    g, sends, counters, epoch = peer.begin_fused(ctx, op)
    op.apply_overlapped(dst.data_ptr(), src.data_ptr(), 0.0, g.data_ptr(), sends, counters, epoch)
    peer.consumed(ctx)
timeit("fused halo kernel step", fused)
timeit("serial peer step", serial)
dist.barrier(); dist.destroy_process_group()
