"""Phase-space partition across the GPUs of one box and the ghost-face exchange between the bricks.

Host-side plumbing only (torch.distributed: NCCL on GPUs, gloo in the CPU tests); the data movement on the device
is done by libhdgpu (hd_halo_pack[_ex], hd_advection_apply[_part]).

Reference counterparts:
  * process grid PartitionX x PartitionV of the drivers (performance/util/driver.h:79-126,
    examples/advection/performance/weak.py:95-101 doubles the x-directions first) -> ``BrickPartition``:
    a Cartesian brick decomposition with one entry per direction;
  * VectorDataExchange::Contiguous::export_to_ghosted_array_start/finish
    (matrix_free/vector_partitioner.h:1387-1592): pack, MPI_Isend/Irecv per neighbour, wait -> ``HaloExchange``:
    one grouped batch of send/recv pairs per operator application, started before the interior cells and
    finished before the boundary-layer cells.
"""
from __future__ import annotations

from dataclasses import dataclass

SIDE_PERIODIC_LOCAL, SIDE_GHOST = 0, 1


class BrickPartition:
    """world ranks -> p[0] x ... x p[dim-1] grid of bricks; rank = lexicographic index, direction 0 fastest."""

    def __init__(self, world: int, rank: int, n_cells_local, split_order=(2, 1, 0), grid=None):
        self.world, self.rank = int(world), int(rank)
        self.dim = len(n_cells_local)
        self.n_cells = tuple(int(c) for c in n_cells_local)
        if grid is None:
            # powers of two are spread over split_order round-robin (weak scaling: x_2, x_1, x_0, x_2, ...)
            grid = [1] * self.dim
            w, i = self.world, 0
            while w > 1:
                if w % 2:
                    raise ValueError("world size must be a power of two unless an explicit grid is given")
                grid[split_order[i % len(split_order)]] *= 2
                w //= 2
                i += 1
        self.grid = tuple(int(g) for g in grid)
        n = 1
        for g in self.grid:
            n *= g
        if n != self.world:
            raise ValueError("grid %r does not match world size %d" % (self.grid, self.world))
        self.coords = self.coords_of(self.rank)
        self.n_cells_global = tuple(c * g for c, g in zip(self.n_cells, self.grid))
        self.cell_offset = tuple(c * k for c, k in zip(self.n_cells, self.coords))
        # a direction cut into several bricks has ghost sides; an uncut periodic direction wraps inside the brick
        self.side_kind = [[SIDE_GHOST if g > 1 else SIDE_PERIODIC_LOCAL] * 2 for g in self.grid]

    def coords_of(self, rank: int):
        c, r = [], rank
        for g in self.grid:
            c.append(r % g)
            r //= g
        return tuple(c)

    def rank_of(self, coords) -> int:
        r, m = 0, 1
        for c, g in zip(coords, self.grid):
            r += (c % g) * m
            m *= g
        return r

    def neighbour(self, d: int, side: int) -> int:
        """rank owning the brick behind side (d, side) (periodic wrap)."""
        c = list(self.coords)
        c[d] += 1 if side else -1
        return self.rank_of(c)


def ghost_layout(n_cells_local, dofs_1d: int, side_kind):
    """Offsets/sizes (in values) of the ghost segments, ordered (direction, side) — the layout of hd_halo_offset /
    hd_mesh_ghost_size (include/hyperdeal_b200.h), i.e. n_face_cells * (k+1)^(dim-1) values per ghost side."""
    dim = len(n_cells_local)
    ncells = 1
    for c in n_cells_local:
        ncells *= c
    nf = dofs_1d ** (dim - 1)
    off, sizes, offsets = 0, {}, {}
    for d in range(dim):
        for s in range(2):
            offsets[(d, s)] = off
            if side_kind[d][s] == SIDE_GHOST:
                sizes[(d, s)] = (ncells // n_cells_local[d]) * nf
                off += sizes[(d, s)]
            else:
                sizes[(d, s)] = 0
    return offsets, sizes, off


def ctypes_ptr(p):
    import ctypes

    return ctypes.c_void_p(int(p))


@dataclass
class _Msg:
    d: int
    side: int  # my side the message belongs to (send: my boundary layer; recv: my ghost side)
    peer: int
    offset: int
    size: int


class HaloExchange:
    """Ghost-face exchange plan of one brick.

    needed[2*d+side] (optional) = the operator reads ghost side (d, side); with the upwind flux only the inflow side of
    every direction is read (hd_advection_ghost_sides), which halves the traffic.  All ranks must pass the same mask.
    """

    def __init__(self, part: BrickPartition, offsets, sizes, needed=None):
        self.part = part
        self.sends, self.recvs = [], []
        for d in range(part.dim):
            if part.grid[d] == 1:
                continue
            # my boundary layer `side` fills the ghost side (1 - side) of the neighbour behind it
            for side in (0, 1):
                if needed is None or needed[2 * d + (1 - side)]:
                    self.sends.append(_Msg(d, side, part.neighbour(d, side), offsets[(d, side)], sizes[(d, side)]))
            # Receives are posted upper side first: NCCL matches the messages of one pair of ranks by order, and when
            # a direction is cut in two the same peer is both neighbours — its first send (its lower layer) is my
            # upper ghost.
            for side in (1, 0):
                if needed is None or needed[2 * d + side]:
                    self.recvs.append(_Msg(d, side, part.neighbour(d, side), offsets[(d, side)], sizes[(d, side)]))

    def send_mask(self, max_dim: int = 6):
        m = [0] * (2 * max_dim)
        for s in self.sends:
            m[2 * s.d + s.side] = 1
        return m

    @property
    def bytes_per_exchange(self):
        return sum(s.size for s in self.sends), sum(r.size for r in self.recvs)

    def start(self, send, ghost, group=None):
        """Post all sends/receives (send, ghost: 1-D tensors holding the packed layers / the ghost segments)."""
        import torch.distributed as dist

        ops = []
        for s in self.sends:
            ops.append(dist.P2POp(dist.isend, send[s.offset : s.offset + s.size], s.peer, group=group, tag=2 * s.d + s.side))
        for r in self.recvs:
            # the sender tagged the message with ITS side = the opposite of my ghost side
            ops.append(dist.P2POp(dist.irecv, ghost[r.offset : r.offset + r.size], r.peer, group=group, tag=2 * r.d + (1 - r.side)))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()


class PeerHaloExchange:
    """Direct variant for one NVLink/NVSwitch box: boundary layers are stored straight into the neighbour GPU's ghost
    segment through peer-mapped pointers — the B200 counterpart of the reference's MPI-3 shared-memory window, where the
    ranks of one node read each other's vector directly (matrix_free/vector_partitioner.h:552-640 `sync`, :1387-1460).

    Two ways to drive it:
      * fused (``begin_fused`` + AdvectionOperation.apply_overlapped): the first ``op.n_halo_senders`` CTAs of the operator
        kernel start with the packing and the NVLink stores and bump the receiver's arrival counter, then join the others
        on the cells; the kernel's boundary phase waits for its own counters.  One launch per operator application, no extra kernels.  (A separate pack kernel cannot run
        beside the persistent operator kernel once that one is resident — the SM sub-partitions' register files are
        full — so "pack on a side stream" only overlaps when it wins the launch race.)
      * split (``start`` / ``wait_ready``): hd_halo_pack_ex as its own kernel, data-ready flags written and awaited with
        stream memory operations, operator in two parts (interior, boundary).
    In both, ghost buffers are double-buffered (step m uses buffer m % 2) and a backward "consumed" flag per message
    (``consumed``) tells the sender when a buffer may be overwritten: before step m it waits for ack >= m - 2.
    Buffers, counters and flags are symmetric-memory allocations (torch.distributed._symmetric_memory provides the
    peer-mapped addresses); all synchronisation is device-side (stream memory operations, no kernels, no host waits).
    Requires equal bricks on all ranks (same ghost layout and CTA count everywhere).
    """

    COUNT, READY, ACK = 0, 16, 32  # word offsets of the three flag groups (slot = 2*d+side inside each)

    def __init__(self, part: BrickPartition, offsets, sizes, total: int, needed, device, group=None, max_dim: int = 6, dtype=None, transport=None, ctx=None):
        """transport: "symm" = torch.distributed symmetric memory (the default), "ipc" = plain device allocations exchanged
        through CUDA IPC handles (hd_device_malloc / hd_ipc_export / hd_ipc_open; needs `ctx`, the api.Context of this
        rank) — what an MPI host would do; HD_PEER_TRANSPORT overrides the default."""
        import os

        import torch
        import torch.distributed as dist

        self.part = part
        group = group if group is not None else dist.group.WORLD
        n = max(int(total), 16)
        dtype = torch.float64 if dtype is None else dtype  # the lattice's number type (hd_mesh_desc.number_type)
        esize = torch.empty(0, dtype=dtype).element_size()
        transport = transport or os.environ.get("HD_PEER_TRANSPORT", "symm")
        if transport == "ipc" and ctx is None:
            raise ValueError("the ipc transport needs the api.Context of this rank")
        self.transport = transport
        self.plan = HaloExchange(part, offsets, sizes, needed)
        if transport == "ipc":
            ghost_ptrs, self.flag_ptrs = self._ipc_setup(ctx, n * esize, group)
        else:
            import torch.distributed._symmetric_memory as symm

            self.ghosts = [symm.empty(n, dtype=dtype, device=device) for _ in range(2)]
            self.flags = symm.empty(64, dtype=torch.int32, device=device)
            for g in self.ghosts:
                g.zero_()
            self.flags.zero_()
            torch.cuda.synchronize()
            self.handles = [symm.rendezvous(g, group) for g in self.ghosts]
            self.flag_handle = symm.rendezvous(self.flags, group)
            ghost_ptrs = [[int(x) for x in h.buffer_ptrs] for h in self.handles]
            self.flag_ptrs = [int(x) for x in self.flag_handle.buffer_ptrs]
        self.mask = self.plan.send_mask(max_dim)
        self.bytes_sent = self.plan.bytes_per_exchange[0] * esize
        self.my_flags = self.flag_ptrs[part.rank]
        self.peer_dst, self.fused_sends = [], []
        for bptrs in ghost_ptrs:
            ptrs, sends = [0] * (2 * max_dim), []
            for s in self.plan.sends:
                # my boundary layer (d, side) is the ghost segment (d, 1 - side) of the neighbour behind that side
                ptrs[2 * s.d + s.side] = bptrs[s.peer] + esize * offsets[(s.d, 1 - s.side)]
                sends.append((s.d, s.side, ptrs[2 * s.d + s.side], self.flag_ptrs[s.peer] + 4 * (self.COUNT + 2 * s.d + (1 - s.side))))
            self.peer_dst.append(ptrs)
            self.fused_sends.append(sends)
        self.step = 0
        self.fused_steps = 0
        dist.barrier(group=group)  # every rank has zeroed its flags before anyone signals

    class _Raw:
        """a device allocation that is not a torch tensor: the callers only ask for data_ptr()"""

        def __init__(self, ptr):
            self.ptr = int(ptr)

        def data_ptr(self):
            return self.ptr

    def _ipc_setup(self, ctx, ghost_bytes, group):
        """two ghost buffers + one flag block per rank from hd_device_malloc; handles all-gathered; the allocations of the
        ranks this brick talks to are mapped with hd_ipc_open.  Returns ([ptrs of buffer 0 by rank, ptrs of buffer 1], flag ptrs)."""
        import ctypes

        import torch.distributed as dist

        from . import api

        L = api.lib()
        self._ctx, self._own, self._opened = ctx, [], []
        mine = []
        for nbytes in (ghost_bytes, ghost_bytes, 256):
            p = ctypes.c_void_p()
            api._check(L.hd_device_malloc(ctx._h, int(nbytes), ctypes.byref(p)))
            h = ctypes.create_string_buffer(64)
            api._check(L.hd_ipc_export(ctx._h, p, h))
            self._own.append(p.value)
            mine.append(bytes(h.raw))
        everyone = [None] * self.part.world
        dist.all_gather_object(everyone, mine, group=group)
        peers = {m.peer for m in self.plan.sends} | {m.peer for m in self.plan.recvs}
        table = [[0] * self.part.world for _ in range(3)]
        for r in range(self.part.world):
            for k in range(3):
                if r == self.part.rank:
                    table[k][r] = self._own[k]
                elif r in peers:
                    p = ctypes.c_void_p()
                    api._check(L.hd_ipc_open(ctx._h, everyone[r][k], ctypes.byref(p)))
                    self._opened.append(p.value)
                    table[k][r] = p.value
        self.ghosts = [PeerHaloExchange._Raw(self._own[0]), PeerHaloExchange._Raw(self._own[1])]
        return [table[0], table[1]], table[2]

    def close(self):
        """ipc transport: unmap the peers' allocations and free this rank's (call after a barrier: nobody may still write)"""
        if getattr(self, "_own", None):
            from . import api

            L = api.lib()
            for p in self._opened:
                L.hd_ipc_close(self._ctx._h, ctypes_ptr(p))
            for p in self._own:
                L.hd_device_free(self._ctx._h, ctypes_ptr(p))
            self._own, self._opened = [], []

    def _next(self, ctx):
        self.step += 1
        m = self.step
        if m > 2:
            for s in self.plan.sends:
                ctx.wait_flag(self.my_flags + 4 * (self.ACK + 2 * s.d + s.side), m - 2)
        return m

    # ---- fused: pack + transport inside the operator kernel
    def begin_fused(self, ctx, op):
        """Enqueue the buffer hand-shake on ctx's stream; returns (ghost tensor, sends, counters_ptr, target) for
        AdvectionOperation.apply_overlapped, to be followed by ``consumed``.  Every sender CTA of a neighbour adds 1 to my arrival
        counter per application, hence target = applications * op.n_halo_senders (equal bricks on all ranks)."""
        m = self._next(ctx)
        self.fused_steps += 1
        return self.ghosts[m % 2], self.fused_sends[m % 2], self.my_flags + 4 * self.COUNT, self.fused_steps * op.n_halo_senders

    # ---- split: pack kernel + stream flags + two operator launches
    def start(self, mf, ctx, src_ptr: int):
        """Enqueue on ctx's current stream: wait for the buffers to be free, pack-and-store into the neighbours, signal.
        Returns (ghost tensor, step) for the operator parts of this step."""
        m = self._next(ctx)
        mf.halo_pack(src_ptr, None, send_mask=self.mask, peer_dst=self.peer_dst[m % 2])
        for s in self.plan.sends:
            ctx.write_flag(self.flag_ptrs[s.peer] + 4 * (self.READY + 2 * s.d + (1 - s.side)), m)
        return self.ghosts[m % 2], m

    def wait_ready(self, ctx, step: int):
        """make ctx's current stream wait until this step's ghost faces have arrived"""
        for r in self.plan.recvs:
            ctx.wait_flag(self.my_flags + 4 * (self.READY + 2 * r.d + r.side), step)

    def consumed(self, ctx):
        """enqueue behind the operator: tell the senders that the ghost buffer of this step may be overwritten"""
        for r in self.plan.recvs:
            # the sender's layer was its side (1 - r.side)
            ctx.write_flag(self.flag_ptrs[r.peer] + 4 * (self.ACK + 2 * r.d + (1 - r.side)), self.step)
