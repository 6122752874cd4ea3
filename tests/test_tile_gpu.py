"""The tile kernel (kernels_generic.cu: k_apply_tile — degree 3, 1D1V / 2D2V, periodic and ghost sides; BASELINE.json
configs[0] runs on it) against the oracle's literal ECL algorithm, on the same seeded inputs."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

VEL = np.array([1.0, 0.15, -0.05, 0.1, -0.15, 0.5])


@pytest.fixture(scope="module")
def api():
    from hyperdeal_b200 import api as A

    return A


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _run(api, ctx, dx, dv, nc, nq=None, colloc=False, skew=0.0, dtype=np.float64, vel=None, kernel=3, seed=5):
    dim, k = dx + dv, 3
    vel = VEL[:dim] if vel is None else np.asarray(vel, dtype=np.float64)
    left, right = (-1.0,) * dim, (1.0,) * dim
    om = O.Mesh(dx, dv, tuple(nc), left, right, (True,) * dim)
    orc = O.Oracle(om, k, nq=nq, collocation=colloc, skew=skew, velocity=vel, nthreads=8)
    src = np.random.default_rng(seed).standard_normal(orc.ndofs)
    if dtype == np.float32:
        src = src.astype(np.float32).astype(np.float64)
    ref = orc.apply(src, time=0.0)
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right, n_points=nq, collocation=colloc, dtype=dtype)
    op = api.AdvectionOperation(mf, vel, skew)
    op.set_kernel(kernel)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.0)
    out = mf.copy_out(d_dst).astype(np.float64)
    name = op.kernel_name
    for p in (d_src, d_dst):
        mf.free_vector(p)
    op.close()
    mf.close()
    return _rel(out, ref), name


CASES = [
    # dx dv cells            nq    colloc skew  velocity
    (1, 1, (4, 3), None, False, 0.0, None),
    (1, 1, (1, 1), None, False, 0.5, None),                 # one cell: both neighbours are the cell itself
    (1, 1, (37, 11), None, False, 0.5, None),               # 407 cells: two CTAs, the second one partly filled
    (1, 1, (300, 2), None, False, 0.5, (-1.0, 0.4)),        # direction-0 neighbours across CTA boundaries, upper-side upwind
    (2, 2, (3, 2, 4, 2), None, False, 0.0, None),
    (2, 2, (3, 2, 4, 2), None, False, 0.5, None),
    (2, 2, (5, 3, 4, 3), None, False, 0.5, None),           # 180 cells: 12 CTAs, the last one partly filled
    (2, 2, (3, 2, 4, 2), 5, False, 0.5, None),              # over-integration
    (2, 2, (2, 2, 2, 2), None, True, 0.3, None),            # collocation
    (2, 2, (1, 1, 1, 1), None, False, 1.0, None),
    (2, 2, (17, 2, 1, 3), None, False, 0.5, (-1.0, -0.15, 0.05, -0.1)),  # all signs flipped
    (2, 2, (4, 4, 2, 2), None, False, 0.5, (0.0, 0.3, 0.0, -0.2)),       # zero components: no traces in those directions
    (3, 3, (2, 2, 2, 2, 2, 2), None, False, 0.5, None),     # three rounds, one cell per CTA
    (3, 3, (3, 2, 1, 2, 2, 3), None, False, 0.0, None),     # ragged, a direction with one cell
    (3, 3, (4, 1, 2, 3, 1, 2), None, False, 0.5, (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5)),
    (3, 3, (2, 3, 2, 1, 2, 2), None, False, 0.5, (0.0, 0.3, 0.0, 0.0, -0.2, 0.0)),
]


@pytest.mark.parametrize("dx,dv,nc,nq,colloc,skew,vel", CASES)
def test_tile_kernel_matches_oracle_f64(api, ctx, dx, dv, nc, nq, colloc, skew, vel):
    rel, name = _run(api, ctx, dx, dv, nc, nq, colloc, skew, vel=vel)
    assert name == "tile"
    assert rel <= 1e-12, rel


@pytest.mark.parametrize("dx,dv,nc", [(1, 1, (9, 5)), (2, 2, (3, 2, 2, 3)), (3, 3, (2, 2, 1, 2, 3, 2))])
def test_tile_kernel_float(api, ctx, dx, dv, nc):
    rel, name = _run(api, ctx, dx, dv, nc, skew=0.5, dtype=np.float32)
    assert name == "tile" and rel <= 1e-5, rel


def test_auto_selects_tile_kernel_for_3d3v_float(api, ctx):
    """FP64 3D3V goes to the pipelined kernel, FP32 3D3V to the tile kernel"""
    rel, name = _run(api, ctx, 3, 3, (2, 2, 2, 2, 2, 2), skew=0.5, kernel=0)
    assert name in ("advect_3d3v_k3", "rounds_3d3v_k3") and rel <= 1e-12
    rel, name = _run(api, ctx, 3, 3, (2, 2, 2, 2, 2, 2), skew=0.5, kernel=0, dtype=np.float32)
    assert name == "tile" and rel <= 1e-5


def test_auto_selects_tile_kernel_and_agrees_with_generic(api, ctx):
    rel, name = _run(api, ctx, 2, 2, (3, 2, 4, 2), skew=0.5, kernel=0)
    assert name == "tile" and rel <= 1e-12
    rel, name = _run(api, ctx, 2, 2, (3, 2, 4, 2), skew=0.5, kernel=1)
    assert name == "generic" and rel <= 1e-12


def test_tile_kernel_refuses_what_it_does_not_cover(api, ctx):
    mf = api.MatrixFree(ctx, 2, 1, 3, (2,) * 3, (0.0,) * 3, (1.0,) * 3)  # odd number of directions
    op = api.AdvectionOperation(mf, VEL[:3], 0.5)
    with pytest.raises(api.HdError):
        op.set_kernel(3)
    mf2 = api.MatrixFree(ctx, 2, 2, 3, (2,) * 4, (0.0,) * 4, (1.0,) * 4, periodic=False)
    op2 = api.AdvectionOperation(mf2, VEL[:4], 0.5)
    with pytest.raises(api.HdError):
        op2.set_kernel(3)
    mf3 = api.MatrixFree(ctx, 2, 2, 2, (2,) * 4, (0.0,) * 4, (1.0,) * 4)
    op3 = api.AdvectionOperation(mf3, VEL[:4], 0.5)
    with pytest.raises(api.HdError):
        op3.set_kernel(3)


@pytest.mark.parametrize("dx,dv,split_dir", [(2, 2, 0), (2, 2, 1), (2, 2, 2), (2, 2, 3), (1, 1, 0), (1, 1, 1), (3, 3, 0), (3, 3, 3), (3, 3, 5)])
@pytest.mark.parametrize("vel_sign", [1.0, -1.0])
def test_tile_kernel_two_bricks_with_ghost_faces(api, ctx, dx, dv, split_dir, vel_sign):
    """two bricks along one direction, faces exchanged by hand through hd_halo_pack: the ghost path of the tile kernel"""
    dim, k = dx + dv, 3
    vel = vel_sign * VEL[:dim]
    nc = [3, 2, 2, 3, 2, 2][:dim]
    nc[split_dir] = 4
    left, right = (-1.0,) * dim, (1.0,) * dim
    om = O.Mesh(dx, dv, tuple(nc), left, right, (True,) * dim)
    orc = O.Oracle(om, k, skew=0.5, velocity=vel, nthreads=8)
    src = np.random.default_rng(9).standard_normal(orc.ndofs)
    ref = orc.apply(src)
    nd = 4**dim
    full = src.reshape(tuple(reversed(nc)) + (nd,))
    ref_full = ref.reshape(tuple(reversed(nc)) + (nd,))
    axis = dim - 1 - split_dir
    bricks = []
    for b in range(2):
        loc = list(nc)
        loc[split_dir] = 2
        off = [0] * dim
        off[split_dir] = 2 * b
        side_kind = [[api.SIDE_PERIODIC_LOCAL] * 2 for _ in range(dim)]
        side_kind[split_dir] = [api.SIDE_GHOST, api.SIDE_GHOST]
        mf = api.MatrixFree(ctx, dx, dv, k, loc, left, right, n_cells_global=nc, cell_offset=off, side_kind=side_kind)
        sl = [slice(None)] * (dim + 1)
        sl[axis] = slice(2 * b, 2 * b + 2)
        u = np.ascontiguousarray(full[tuple(sl)]).reshape(-1)
        d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.copy_in(d_src, u)
        d_send, d_ghost = mf.initialize_dof_vector(), mf.initialize_dof_vector()  # big enough
        mf.halo_pack(d_src, d_send)
        send = mf.copy_out(d_send, mf.halo_total)
        bricks.append(dict(mf=mf, src=d_src, dst=d_dst, send=send, ghost=d_ghost, sl=tuple(sl)))
    for b in range(2):
        me, other = bricks[b], bricks[1 - b]
        mf = me["mf"]
        ghost = np.zeros(mf.halo_total)
        for side in range(2):
            o_me, n_me = mf.halo_offset(split_dir, side), mf.ghost_size(split_dir, side)
            o_ot = other["mf"].halo_offset(split_dir, 1 - side)
            ghost[o_me : o_me + n_me] = other["send"][o_ot : o_ot + n_me]
        op = api.AdvectionOperation(mf, vel, 0.5)
        op.set_kernel(3)
        needed = op.ghost_sides()
        for side in range(2):
            if not needed[2 * split_dir + side]:  # the outflow side is never read: poison it
                o_me, n_me = mf.halo_offset(split_dir, side), mf.ghost_size(split_dir, side)
                ghost[o_me : o_me + n_me] = np.nan
        mf.copy_in(me["ghost"], ghost)
        op.apply(me["dst"], me["src"], 0.0, ghosts=me["ghost"])
        assert op.kernel_name == "tile"
        out = mf.copy_out(me["dst"])
        expect = np.ascontiguousarray(ref_full[me["sl"]]).reshape(-1)
        assert _rel(out, expect) <= 1e-12


@pytest.mark.parametrize("rk", ["rk45", "rk33"])
def test_tile_kernel_fused_lsrk(api, ctx, rk):
    dx, dv, nc, k = 2, 2, (5, 2, 2, 3), 3
    left, right = (-1.0,) * 4, (1.0,) * 4
    om = O.Mesh(dx, dv, nc, left, right, (True,) * 4)
    orc = O.Oracle(om, k, skew=0.5, velocity=VEL[:4], nthreads=4)
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right)
    op = api.AdvectionOperation(mf, VEL[:4], 0.5)
    op.set_kernel(3)
    sol0 = np.random.default_rng(1).standard_normal(mf.n_dofs)
    dt, ref = 0.004, sol0
    for s in range(2):
        ref = O.lsrk_step(lambda v, tt: orc.apply(v, tt), ref, s * dt, dt, rk)
    sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(sol, sol0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, rk)
    for s in range(2):
        integ.perform_time_step(sol, s * dt, dt, op)
    assert op.kernel_name == "tile"
    assert _rel(mf.copy_out(sol), ref) <= 1e-12


# ------------------------------------------------------------------------------------------
# the row-persistent 3D3V variant (k_apply_tile_row, hd_advection_set_kernel(4))
ROW_CASES = [
    # cells                 velocity                                   skew
    ((2, 2, 2, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5),
    ((3, 2, 1, 2, 2, 3), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.0),     # ragged, odd row length
    ((4, 1, 2, 3, 1, 2), (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), 0.5),   # descending walk
    ((2, 3, 2, 1, 2, 2), (0.0, 0.3, 0.0, 0.0, -0.2, 0.0), 0.5),         # no direction-0 neighbour at all
    ((1, 2, 2, 2, 1, 2), (0.4, -0.3, 0.2, -0.1, 0.6, 0.7), 1.0),        # rows of one cell: the neighbour is the cell itself
    ((7, 2, 2, 1, 2, 1), (-0.4, 0.3, 0.2, 0.1, 0.6, -0.7), 0.5),
    ((8, 4, 4, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5),     # 512 rows: every CTA walks several rows
]


@pytest.mark.parametrize("nc,vel,skew", ROW_CASES)
def test_tile_row_kernel_matches_oracle(api, ctx, nc, vel, skew):
    rel, name = _run(api, ctx, 3, 3, nc, skew=skew, vel=vel, kernel=4)
    assert name == "tile_row"
    assert rel <= 1e-12, rel


def test_tile_row_kernel_float(api, ctx):
    rel, name = _run(api, ctx, 3, 3, (3, 2, 1, 2, 3, 2), skew=0.5, dtype=np.float32, kernel=4)
    assert name == "tile_row" and rel <= 1e-5, rel


def test_tile_row_kernel_refuses_other_dimensions(api, ctx):
    mf = api.MatrixFree(ctx, 2, 2, 3, (2,) * 4, (0.0,) * 4, (1.0,) * 4)
    op = api.AdvectionOperation(mf, VEL[:4], 0.5)
    with pytest.raises(api.HdError):
        op.set_kernel(4)


@pytest.mark.parametrize("split_dir", [0, 2, 5])
@pytest.mark.parametrize("vel_sign", [1.0, -1.0])
def test_tile_row_kernel_two_bricks_with_ghost_faces(api, ctx, split_dir, vel_sign):
    dx, dv, dim, k = 3, 3, 6, 3
    vel = vel_sign * VEL
    nc = [3, 2, 2, 3, 2, 2]
    nc[split_dir] = 4
    left, right = (-1.0,) * dim, (1.0,) * dim
    om = O.Mesh(dx, dv, tuple(nc), left, right, (True,) * dim)
    orc = O.Oracle(om, k, skew=0.5, velocity=vel, nthreads=8)
    src = np.random.default_rng(9).standard_normal(orc.ndofs)
    ref = orc.apply(src)
    nd = 4**dim
    full = src.reshape(tuple(reversed(nc)) + (nd,))
    ref_full = ref.reshape(tuple(reversed(nc)) + (nd,))
    axis = dim - 1 - split_dir
    bricks = []
    for b in range(2):
        loc = list(nc)
        loc[split_dir] = 2
        off = [0] * dim
        off[split_dir] = 2 * b
        side_kind = [[api.SIDE_PERIODIC_LOCAL] * 2 for _ in range(dim)]
        side_kind[split_dir] = [api.SIDE_GHOST, api.SIDE_GHOST]
        mf = api.MatrixFree(ctx, dx, dv, k, loc, left, right, n_cells_global=nc, cell_offset=off, side_kind=side_kind)
        sl = [slice(None)] * (dim + 1)
        sl[axis] = slice(2 * b, 2 * b + 2)
        u = np.ascontiguousarray(full[tuple(sl)]).reshape(-1)
        d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.copy_in(d_src, u)
        d_send, d_ghost = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.halo_pack(d_src, d_send)
        send = mf.copy_out(d_send, mf.halo_total)
        bricks.append(dict(mf=mf, src=d_src, dst=d_dst, send=send, ghost=d_ghost, sl=tuple(sl)))
    for b in range(2):
        me, other = bricks[b], bricks[1 - b]
        mf = me["mf"]
        ghost = np.zeros(mf.halo_total)
        for side in range(2):
            o_me, n_me = mf.halo_offset(split_dir, side), mf.ghost_size(split_dir, side)
            o_ot = other["mf"].halo_offset(split_dir, 1 - side)
            ghost[o_me : o_me + n_me] = other["send"][o_ot : o_ot + n_me]
        op = api.AdvectionOperation(mf, vel, 0.5)
        op.set_kernel(4)
        needed = op.ghost_sides()
        for side in range(2):
            if not needed[2 * split_dir + side]:
                o_me, n_me = mf.halo_offset(split_dir, side), mf.ghost_size(split_dir, side)
                ghost[o_me : o_me + n_me] = np.nan
        mf.copy_in(me["ghost"], ghost)
        op.apply(me["dst"], me["src"], 0.0, ghosts=me["ghost"])
        assert op.kernel_name == "tile_row"
        out = mf.copy_out(me["dst"])
        expect = np.ascontiguousarray(ref_full[me["sl"]]).reshape(-1)
        assert _rel(out, expect) <= 1e-12


def test_tile_row_kernel_fused_lsrk(api, ctx):
    dx, dv, nc, k = 3, 3, (3, 2, 2, 1, 2, 2), 3
    left, right = (-1.0,) * 6, (1.0,) * 6
    om = O.Mesh(dx, dv, nc, left, right, (True,) * 6)
    orc = O.Oracle(om, k, skew=0.5, velocity=VEL, nthreads=4)
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right)
    op = api.AdvectionOperation(mf, VEL, 0.5)
    op.set_kernel(4)
    sol0 = np.random.default_rng(1).standard_normal(mf.n_dofs)
    dt, ref = 0.004, sol0
    for s in range(2):
        ref = O.lsrk_step(lambda v, tt: orc.apply(v, tt), ref, s * dt, dt, "rk45")
    sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(sol, sol0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk45")
    for s in range(2):
        integ.perform_time_step(sol, s * dt, dt, op)
    assert op.kernel_name == "tile_row"
    assert _rel(mf.copy_out(sol), ref) <= 1e-12
