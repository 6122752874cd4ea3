"""Host-side multi-GPU logic on CPU: brick partition + ghost-face exchange over torch.distributed (gloo),
world sizes 2 and 4.  The device pack kernel is replaced here by a numpy restatement of the same layout
(tests only); what is checked is the plumbing of hyperdeal_b200/partition.py: neighbour ranks, segment
offsets, message matching (including the cut-in-two case where both neighbours are the same rank) and the
upwind-only mask."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperdeal_b200.partition import BrickPartition, HaloExchange, ghost_layout

N1D = 2  # degree 1: small faces


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _field(ncg, dim):
    """global field u[cell coords (reversed), dof coords (reversed)] with a unique value per (cell, dof)."""
    shape = tuple(reversed(ncg)) + (N1D,) * dim
    return np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape)


def _layer(block, dim, d, side):
    """nodal face layer of direction d of every boundary cell of `block` (cells reversed + dofs reversed axes),
    ordered face-cell major, then face dofs — the order hd_halo_pack writes."""
    cell_axis = dim - 1 - d
    dof_axis = dim + (dim - 1 - d)
    sl = [slice(None)] * (2 * dim)
    sl[cell_axis] = -1 if side else 0
    sl[dof_axis] = -1 if side else 0
    return np.ascontiguousarray(block[tuple(sl)]).reshape(-1)


def _worker(rank, world, port, nloc, split_order, needed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = len(nloc)
        part = BrickPartition(world, rank, nloc, split_order=split_order)
        u = _field(part.n_cells_global, dim)
        sl = tuple(slice(part.cell_offset[d], part.cell_offset[d] + nloc[d]) for d in reversed(range(dim)))
        mine = u[sl]
        offsets, sizes, total = ghost_layout(nloc, N1D, part.side_kind)
        send = np.full(total, -1.0)
        for d in range(dim):
            for s in range(2):
                if sizes[(d, s)]:
                    send[offsets[(d, s)] : offsets[(d, s)] + sizes[(d, s)]] = _layer(mine, dim, d, s)
        ex = HaloExchange(part, offsets, sizes, needed)
        t_send, t_ghost = torch.from_numpy(send), torch.full((total,), -7.0, dtype=torch.float64)
        HaloExchange.finish(ex.start(t_send, t_ghost))
        ghost = t_ghost.numpy()
        ok = True
        for d in range(dim):
            for s in range(2):
                n = sizes[(d, s)]
                got = ghost[offsets[(d, s)] : offsets[(d, s)] + n]
                if n == 0:
                    continue
                if needed is not None and not needed[2 * d + s]:
                    ok &= bool(np.all(got == -7.0))  # untouched
                    continue
                # expected: the layer (1 - s) of the brick behind my side s (periodic wrap over the global lattice)
                nb_off = list(part.cell_offset)
                nb_off[d] = (nb_off[d] + (nloc[d] if s else -nloc[d])) % part.n_cells_global[d]
                nsl = tuple(slice(nb_off[e], nb_off[e] + nloc[e]) for e in reversed(range(dim)))
                ok &= bool(np.array_equal(got, _layer(u[nsl], dim, d, 1 - s)))
                ok &= part.neighbour(d, s) == part.rank_of([(part.coords[e] + ((1 if s else -1) if e == d else 0)) for e in range(dim)])
        q.put((rank, ok, part.grid))
    finally:
        dist.destroy_process_group()


def _run(world, nloc, split_order, needed=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nloc, split_order, needed, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


@pytest.mark.timeout(180)
def test_two_ranks_cut_in_two():
    """world 2: the only cut direction has the same rank on both sides (message order matters on NCCL)."""
    res = _run(2, (2, 3, 2, 2), split_order=(1, 0))
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == (1, 2, 1, 1)


@pytest.mark.timeout(180)
def test_four_ranks_two_directions_upwind_only():
    needed = [0] * 12
    needed[2 * 2 + 1] = 1  # direction 2: only the upper ghost side is read (a_2 < 0)
    needed[2 * 1 + 0] = 1  # direction 1: only the lower one
    res = _run(4, (2, 2, 3, 2, 1, 2), split_order=(2, 1, 0), needed=needed)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == (1, 2, 2, 1, 1, 1)


def test_partition_grids_follow_the_weak_scaling_recipe():
    # examples/advection/performance/weak.py:95-101: x-directions are doubled first; here x_2, x_1, x_0
    assert BrickPartition(1, 0, (8,) * 6).grid == (1,) * 6
    assert BrickPartition(2, 1, (8,) * 6).grid == (1, 1, 2, 1, 1, 1)
    assert BrickPartition(4, 3, (8,) * 6).grid == (1, 2, 2, 1, 1, 1)
    p = BrickPartition(8, 5, (8,) * 6)
    assert p.grid == (2, 2, 2, 1, 1, 1) and p.n_cells_global == (16, 16, 16, 8, 8, 8)
    assert p.coords == (1, 0, 1, 0, 0, 0) and p.cell_offset == (8, 0, 8, 0, 0, 0)
    assert p.neighbour(0, 1) == 4 and p.neighbour(2, 0) == 1 and p.neighbour(1, 1) == 7
    assert p.side_kind[0] == [1, 1] and p.side_kind[3] == [0, 0]
    with pytest.raises(ValueError):
        BrickPartition(3, 0, (8,) * 6)
    g = BrickPartition(6, 4, (4, 4, 4), grid=(3, 2, 1))
    assert g.coords == (1, 1, 0) and g.neighbour(0, 0) == 3 and g.neighbour(0, 1) == 5


def test_ghost_layout_matches_header_contract():
    offsets, sizes, total = ghost_layout((2, 3, 4), 4, [[1, 1], [0, 0], [1, 1]])
    assert sizes[(0, 0)] == 3 * 4 * 16 and sizes[(2, 1)] == 2 * 3 * 16 and sizes[(1, 0)] == 0
    assert offsets[(0, 1)] == sizes[(0, 0)] and offsets[(2, 0)] == 2 * sizes[(0, 0)]
    assert total == 2 * sizes[(0, 0)] + 2 * sizes[(2, 0)]


def test_x24_layout_of_the_8_gpu_bench():
    """bench.py --layout x24: the 8-GPU lattice (16,16,16,8,8,8) cut as x_2 in four and x_1 in two; bricks of
    16x8x4x8x8x8 cells, neighbours in a direction cut in four are different ranks on the two sides"""
    recipe = BrickPartition(8, 0, (8,) * 6, split_order=(2, 1, 0))
    for rank in range(8):
        cut = BrickPartition(8, rank, (8,) * 6, split_order=(2, 1, 2))
        assert cut.grid == (1, 2, 4, 1, 1, 1)
        nloc = [g // c for g, c in zip(recipe.n_cells_global, cut.grid)]
        part = BrickPartition(8, rank, nloc, grid=cut.grid)
        assert part.n_cells == (16, 8, 4, 8, 8, 8) and part.n_cells_global == recipe.n_cells_global
        lo, hi = part.neighbour(2, 0), part.neighbour(2, 1)
        assert lo != hi and lo != rank
        # neighbour relations are mutual
        assert BrickPartition(8, lo, nloc, grid=cut.grid).neighbour(2, 1) == rank
        assert BrickPartition(8, hi, nloc, grid=cut.grid).neighbour(2, 0) == rank
        assert part.neighbour(1, 0) == part.neighbour(1, 1)  # cut in two: the same rank on both sides
        assert part.side_kind[0] == [0, 0] and part.side_kind[1] == [1, 1] and part.side_kind[2] == [1, 1]


@pytest.mark.timeout(240)
def test_eight_ranks_x24_exchange():
    """the halo exchange plan on the x24 grid (1,2,4): every rank receives exactly its neighbours' boundary layers"""
    res = _run(8, (2, 2, 1, 1, 1, 1), split_order=(2, 1, 2))
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == (1, 2, 4, 1, 1, 1)
