"""Parity of the CUDA advection operator (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): 1e-12 relative in double, 1e-5 in float, on identical meshes,
degrees and data.  "relative" = max|gpu - oracle| / max|oracle| over the vector.
Inputs: numpy.random.default_rng(20240229).standard_normal DoF vectors (every face carries
data) and the sin*cos wave of examples/advection/cases/hyperrectangle.h.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

VEL = np.array([1.0, 0.15, -0.05, 0.1, -0.15, 0.5])
TOL64 = 1e-12
FAST_NAMES = ("advect_3d3v_k3", "rounds_3d3v_k3")  # the two 3D3V degree-3 FP64 kernels (hd_advection_set_kernel 2 / 6)
TOL32 = 1e-5


@pytest.fixture(scope="module")
def api():
    from hyperdeal_b200 import api as A

    return A


@pytest.fixture(scope="module")
def ctx(api):
    c = api.Context(0)
    yield c
    c.close()


def _mesh_pair(api, ctx, dx, dv, nc, degree, nq=None, colloc=False, periodic=True, dtype=np.float64, left=None, right=None):
    dim = dx + dv
    left = left if left is not None else tuple(-1 + 0.1 * d for d in range(dim))
    right = right if right is not None else tuple(1 + 0.2 * d for d in range(dim))
    per = (periodic,) * dim if isinstance(periodic, bool) else periodic
    om = O.Mesh(dx, dv, tuple(nc), left, right, per)
    mf = api.MatrixFree(ctx, dx, dv, degree, nc, left, right, periodic=per, n_points=nq, collocation=colloc, dtype=dtype)
    return om, mf


def _rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b)) / np.max(np.abs(b)))


def _run(api, ctx, dx, dv, nc, degree, nq=None, colloc=False, skew=0.0, periodic=True, dtype=np.float64, vel=None, bc_kind=1, kernel=0, seed=20240229):
    dim = dx + dv
    vel = VEL[:dim] if vel is None else np.asarray(vel, dtype=np.float64)
    om, mf = _mesh_pair(api, ctx, dx, dv, nc, degree, nq, colloc, periodic, dtype)
    orc = O.Oracle(om, degree, nq=nq, collocation=colloc, skew=skew, velocity=vel, bc_kind=bc_kind, nthreads=8)
    rng = np.random.default_rng(seed)
    src = rng.standard_normal(orc.ndofs)
    if dtype == np.float32:
        src = src.astype(np.float32).astype(np.float64)
    ref = orc.apply(src, time=0.3)
    op = api.AdvectionOperation(mf, vel, skew)
    if not all(om.periodic):
        op.set_dirichlet_builtin(api.FN_HYPERRECTANGLE if bc_kind == 1 else api.FN_ZERO)
    op.set_kernel(kernel)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.3)
    out = mf.copy_out(d_dst)
    name = op.kernel_name
    assert op.launch_count >= 1
    mf.free_vector(d_src)
    mf.free_vector(d_dst)
    op.close()
    mf.close()
    return _rel(out, ref), name


CASES64 = [
    # dx dv cells                  k  nq    colloc skew
    (1, 1, (4, 3), 3, None, False, 0.0),
    (1, 1, (1, 1), 3, None, False, 0.5),  # one cell: both neighbours are the cell itself
    (1, 1, (5, 2), 1, None, False, 0.0),
    (2, 1, (2, 3, 2), 2, None, False, 1.0),
    (2, 2, (3, 2, 4, 2), 3, None, False, 0.0),
    (2, 2, (3, 2, 4, 2), 3, None, False, 0.5),
    (2, 2, (3, 2, 4, 2), 3, 5, False, 0.5),  # over-integration
    (2, 2, (2, 2, 2, 2), 3, None, True, 0.3),  # collocation
    (2, 2, (2, 3, 2, 1), 4, None, False, 0.5),
    (3, 2, (2, 2, 2, 3, 2), 3, None, False, 0.5),
    (3, 3, (2, 2, 2, 2, 2, 2), 3, None, False, 0.0),
    (3, 3, (3, 2, 1, 2, 2, 3), 3, None, False, 0.5),  # ragged
    (3, 3, (2, 1, 2, 1, 2, 1), 2, None, False, 0.5),
]


@pytest.mark.parametrize("dx,dv,nc,k,nq,colloc,skew", CASES64)
def test_generic_kernel_matches_oracle_f64(api, ctx, dx, dv, nc, k, nq, colloc, skew):
    rel, name = _run(api, ctx, dx, dv, nc, k, nq, colloc, skew, kernel=1)
    assert name == "generic"
    assert rel <= TOL64, rel


@pytest.mark.parametrize("vel", [(-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), (0.0, 0.3, 0.0, 0.0, -0.2, 0.0), (1.0, 0.0, 0.0, 0.0, 0.0, 0.0)])
def test_generic_kernel_velocity_signs(api, ctx, vel):
    rel, _ = _run(api, ctx, 3, 3, (2, 2, 2, 2, 2, 2), 3, skew=0.5, vel=vel, kernel=1)
    assert rel <= TOL64, rel


# Dirichlet lattices: kernel 0 = the automatic choice, which serves them on the specialised kernels through synthesised ghost
# traces (-u_face + 2 g, capi.cu: apply_dirichlet_as_ghosts); kernel 1 = the generic kernel with boundary matrices + lifting kernel
@pytest.mark.parametrize("kernel,expect", [(0, "tile"), (1, "generic")])
@pytest.mark.parametrize("bc_kind", [1, 2])
@pytest.mark.parametrize("skew", [0.0, 0.5])
def test_dirichlet_matches_oracle(api, ctx, bc_kind, skew, kernel, expect):
    # the reference's Dirichlet goldens (adv_2D_2D_k3.hyperrectangle_03/07) use this mesh class
    dx, dv, nc = 2, 2, (3, 2, 2, 3)
    om, mf = _mesh_pair(api, ctx, dx, dv, nc, 3, periodic=False)
    mf.close()
    om2 = om
    # inhomogeneous (bc_kind 1) / homogeneous (2)
    from hyperdeal_b200 import api as A

    side = A.SIDE_DIRICHLET if bc_kind == 1 else A.SIDE_DIRICHLET_HOM
    mf = A.MatrixFree(ctx, dx, dv, 3, nc, om.left, om.right, side_kind=[[side, side]] * 4)
    vel = np.array([1.0, -0.15, -0.05, 0.2])
    orc = O.Oracle(om2, 3, skew=skew, velocity=vel, bc_kind=bc_kind, nthreads=4)
    src = np.random.default_rng(7).standard_normal(orc.ndofs)
    ref = orc.apply(src, time=0.37)
    op = A.AdvectionOperation(mf, vel, skew)
    op.set_dirichlet_builtin(A.FN_HYPERRECTANGLE)
    op.set_kernel(kernel)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.37)
    out = mf.copy_out(d_dst)
    assert _rel(out, ref) <= TOL64
    assert op.kernel_name == expect


@pytest.mark.parametrize("dx,dv,nc,k,dtype,periodic,vel,expect", [
    # 3D3V degree 3 FP64, all sides Dirichlet: the three-round kernel (ghost traces in all six directions, x_0 included)
    (3, 3, (3, 2, 2, 2, 2, 2), 3, np.float64, False, None, "rounds_3d3v_k3"),
    (3, 3, (2, 2, 2, 2, 2, 2), 3, np.float64, False, (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), "rounds_3d3v_k3"),
    # mixed: periodic x, Dirichlet v; one direction without transport
    (3, 3, (2, 2, 2, 2, 2, 2), 3, np.float64, (True, True, True, False, False, False), (1.0, 0.15, -0.05, 0.0, -0.15, 0.5), "rounds_3d3v_k3"),
    (3, 3, (2, 2, 2, 2, 2, 2), 3, np.float32, False, None, "tile"),
    (2, 2, (3, 2, 2, 3), 3, np.float32, False, None, "tile"),
    (3, 3, (2, 1, 2, 1, 2, 1), 5, np.float32, False, None, "tile_global"),
    (1, 1, (4, 3), 3, np.float64, False, None, "generic"),   # no specialised kernel preferred in 1D1V: lifting-kernel path
])
def test_dirichlet_on_specialised_kernels(api, ctx, dx, dv, nc, k, dtype, periodic, vel, expect):
    rel, name = _run(api, ctx, dx, dv, nc, k, periodic=periodic, dtype=dtype, vel=vel, skew=0.5)
    assert name == expect
    assert rel <= (TOL64 if dtype == np.float64 else TOL32), rel
    rel1, name1 = _run(api, ctx, dx, dv, nc, k, periodic=periodic, dtype=dtype, vel=vel, skew=0.5, kernel=1)
    assert name1 == "generic" and rel1 <= (TOL64 if dtype == np.float64 else TOL32)


def test_dirichlet_uploaded_values(api, ctx):
    """g supplied by the host at the face quadrature points (the BoundaryDescriptor route)."""
    dx, dv, nc, k = 1, 2, (3, 2, 2), 2
    left, right = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)
    om = O.Mesh(dx, dv, nc, left, right, (False,) * 3)
    vel = np.array([0.7, -0.4, 0.2])
    orc = O.Oracle(om, k, skew=0.0, velocity=vel, bc_kind=1, nthreads=2)
    src = np.random.default_rng(11).standard_normal(orc.ndofs)
    t = 0.21
    ref = orc.apply(src, time=t)
    A = api
    mf = A.MatrixFree(ctx, dx, dv, k, nc, left, right, side_kind=[[A.SIDE_DIRICHLET, A.SIDE_DIRICHLET]] * 3)
    op = A.AdvectionOperation(mf, vel, 0.0)
    xq = mf.basis(1)
    nq = len(xq)
    h = om.h
    for d in range(3):
        others = [e for e in range(3) if e != d]
        for side in range(2):
            vals = []
            nfc = [nc[e] for e in others]
            for c1 in range(nfc[1]):
                for c0 in range(nfc[0]):
                    for q1 in range(nq):
                        for q0 in range(nq):
                            p = np.zeros(3)
                            p[d] = left[d] + h[d] * ((nc[d] - 1 if side else 0) + side)
                            p[others[0]] = left[others[0]] + h[others[0]] * (c0 + xq[q0])
                            p[others[1]] = left[others[1]] + h[others[1]] * (c1 + xq[q1])
                            vals.append(O.hyperrectangle_exact(p[None, :], t)[0])
            op.set_dirichlet_values(d, side, np.array(vals))
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, t)
    assert _rel(mf.copy_out(d_dst), ref) <= TOL64


def test_float_k5_3d3v(api, ctx):
    """BASELINE config 3: 3D x 3D, degree 5, float, periodic (small lattice for the oracle)."""
    rel, _ = _run(api, ctx, 3, 3, (2, 1, 1, 1, 1, 2), 5, skew=0.5, dtype=np.float32, kernel=1)
    assert rel <= TOL32, rel


def test_float_k3_2d2v(api, ctx):
    rel, _ = _run(api, ctx, 2, 2, (3, 2, 2, 3), 3, skew=0.0, dtype=np.float32, kernel=1)
    assert rel <= TOL32, rel


def test_apply_host_matches_device_path(api, ctx):
    om, mf = _mesh_pair(api, ctx, 2, 2, (3, 2, 4, 2), 3)
    op = api.AdvectionOperation(mf, VEL[:4], 0.5)
    src = np.random.default_rng(3).standard_normal(mf.n_dofs)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.0)
    a = mf.copy_out(d_dst)
    b = np.empty_like(src)
    op.apply_host(b, src, 0.0)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("a4,a5", [(-0.15, 0.5), (0.15, -0.5), (0.0, 0.0), (0.15, 0.5)])
@pytest.mark.parametrize("n4,n5", [(2, 3), (3, 5), (4, 3), (2, 2)])
def test_apply_host_pipelined_3d3v(api, ctx, a4, a5, n4, n5):
    """hd_advection_apply_host with the pipelined kernel: layers of the slowest direction are copied in, computed and
    copied out on three streams (n5 >= 3, cut again along direction 4 if n4 >= 3; n5 = 2 takes the serial path).  Bit-identical to the device-resident apply,
    for both upwind orientations of the slowest direction and for a_5 = 0."""
    nc = (3, 2, 2, 2, n4, n5)
    mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
    vel = tuple(VEL[:4]) + (a4, a5)
    op = api.AdvectionOperation(mf, vel, 0.5)
    src = np.random.default_rng(5).standard_normal(mf.n_dofs)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.0)
    a = mf.copy_out(d_dst)
    assert op.kernel_name in FAST_NAMES
    for _ in range(2):  # (second call: staging buffers, streams and events are re-used)
        b = np.full_like(src, np.nan)
        op.apply_host(b, src, 0.0)
        assert np.array_equal(a, b)


def test_linearity_and_constant_state(api, ctx):
    """size-independent properties: A(alpha u + v) = alpha A u + A v; A(const) = 0 on a periodic mesh."""
    om, mf = _mesh_pair(api, ctx, 3, 3, (2, 2, 2, 2, 2, 2), 3)
    op = api.AdvectionOperation(mf, VEL, 0.5)
    rng = np.random.default_rng(5)
    u, v = rng.standard_normal(mf.n_dofs), rng.standard_normal(mf.n_dofs)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()

    def A(x):
        mf.copy_in(d_src, x)
        op.apply(d_dst, d_src, 0.0)
        return mf.copy_out(d_dst)

    Au, Av, Aw = A(u), A(v), A(2.5 * u + v)
    assert np.max(np.abs(Aw - (2.5 * Au + Av))) <= 1e-12 * np.max(np.abs(Aw))
    Ac = A(np.full(mf.n_dofs, 3.0))
    assert np.max(np.abs(Ac)) <= 1e-11


def test_errors_are_reported(api, ctx):
    om, mf = _mesh_pair(api, ctx, 1, 1, (2, 2), 3)
    op = api.AdvectionOperation(mf, VEL[:2], 0.0)
    d = mf.initialize_dof_vector()
    with pytest.raises(api.HdError):
        op.apply(d, d, 0.0)  # aliasing
    with pytest.raises(api.HdError):
        op.set_kernel(2)  # the 3D3V kernel does not cover 1D1V
    with pytest.raises(api.HdError):
        op.set_kernel(6)
    with pytest.raises(api.HdError):
        api.MatrixFree(ctx, 4, 1, 3, (1,) * 5, (0,) * 5, (1,) * 5)


# ------------------------------------------------------------------------------------------
# the pipelined 3D3V k=3 kernel (kernel_fast6d.cu)
FAST_CASES = [
    # cells                 velocity                                 skew
    ((2, 2, 2, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5),
    ((3, 2, 1, 2, 2, 3), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.0),   # ragged, a direction with one cell
    ((4, 1, 2, 3, 1, 2), (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5), 0.5),  # all signs flipped: descending walk
    ((2, 3, 2, 1, 2, 2), (0.0, 0.3, 0.0, 0.0, -0.2, 0.0), 0.5),       # zero components: inactive directions
    ((1, 1, 1, 1, 1, 1), (0.4, -0.3, 0.2, -0.1, 0.6, 0.7), 1.0),      # single cell
    ((5, 2, 2, 2, 2, 2), (0.0, 0.0, 0.0, 0.0, 0.0, 0.9), 0.0),
    ((8, 4, 2, 2, 2, 2), (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5),   # more rows than SMs would take at once
]


# kernel 2 = two-role pipelined kernel (advect_3d3v_k3), 6 = three-round kernel (rounds_3d3v_k3, kernel_rounds6d.cuh)
FAST_KERNELS = {2: "advect_3d3v_k3", 6: "rounds_3d3v_k3"}


@pytest.mark.parametrize("kernel", [2, 6])
@pytest.mark.parametrize("nc,vel,skew", FAST_CASES)
def test_fast_kernel_matches_oracle(api, ctx, nc, vel, skew, kernel):
    rel, name = _run(api, ctx, 3, 3, nc, 3, skew=skew, vel=vel, kernel=kernel)
    assert name == FAST_KERNELS[kernel]
    assert rel <= TOL64, rel


def test_auto_selects_fast_kernel(api, ctx):
    rel, name = _run(api, ctx, 3, 3, (2, 2, 2, 2, 2, 2), 3, skew=0.5, kernel=0)
    assert name in FAST_NAMES and rel <= TOL64


@pytest.mark.parametrize("kernel", [2, 6])
def test_fast_kernel_many_rows(api, ctx, kernel):
    """4^6 cells: every CTA walks several rows, all ring/parity phases wrap many times."""
    rel, name = _run(api, ctx, 3, 3, (4, 4, 4, 4, 4, 4), 3, skew=0.5, kernel=kernel)
    assert name == FAST_KERNELS[kernel] and rel <= TOL64


@pytest.mark.parametrize("kernel", [2, 6])
@pytest.mark.parametrize("tile", [(2, 2, 2, 2, 0), (4, 2, 3, 1, 2), (3, 0, 2, 2, 4), (1, 1, 1, 1, 1), (5, 7, 5, 5, 3)])
def test_fast_kernel_row_tiles_are_order_only(api, ctx, tile, kernel):
    """hd_advection_set_row_tile changes the order in which rows of cells are visited (L2 blocking), never the result:
    bit-identical to the lattice order, also for extents the tile does not divide"""
    nc = (3, 4, 2, 6, 3, 4)
    mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
    op = api.AdvectionOperation(mf, (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5)
    op.set_kernel(kernel)
    src = np.random.default_rng(7).standard_normal(mf.n_dofs)
    d_src, d_a, d_b = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(d_src, src)
    op.set_row_tile((0, 0, 0, 0, 0))
    op.apply(d_a, d_src, 0.0)
    op.set_row_tile(tile)
    op.apply(d_b, d_src, 0.0)
    a, b = mf.copy_out(d_a), mf.copy_out(d_b)
    assert np.array_equal(a, b)
    assert np.abs(a).max() > 0
    for p in (d_src, d_a, d_b):
        mf.free_vector(p)


# AdvectionOperationEvaluationLevel (advection_operation.h:37-42): cell integrals only / faces without the neighbour's trace
@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("dx,dv,nc,k,dtype,kernel,expect", [
    (3, 3, (3, 2, 2, 2, 2, 2), 3, np.float64, 0, "rounds_3d3v_k3"),
    (3, 3, (2, 2, 2, 2, 2, 2), 3, np.float64, 2, "advect_3d3v_k3"),
    (2, 2, (3, 2, 4, 2), 3, np.float64, 0, "tile"),
    (2, 2, (3, 2, 4, 2), 3, np.float64, 1, "generic"),
    (2, 1, (2, 3, 2), 2, np.float64, 0, "generic"),
    (3, 3, (2, 1, 2, 1, 2, 1), 5, np.float32, 0, "tile_global"),
])
def test_evaluation_levels_match_oracle(api, ctx, level, dx, dv, nc, k, dtype, kernel, expect):
    dim = dx + dv
    om, mf = _mesh_pair(api, ctx, dx, dv, nc, k, dtype=dtype)
    vel = VEL[:dim]
    orc = O.Oracle(om, k, skew=0.5, velocity=vel, nthreads=8, eval_level=level)
    src = np.random.default_rng(5).standard_normal(orc.ndofs)
    if dtype == np.float32:
        src = src.astype(np.float32).astype(np.float64)
    ref = orc.apply(src, time=0.0)
    full = O.Oracle(om, k, skew=0.5, velocity=vel, nthreads=8).apply(src, time=0.0)
    assert _rel(ref, full) > 1e-3  # the levels are different operators
    op = api.AdvectionOperation(mf, vel, 0.5)
    op.set_kernel(kernel)
    op.set_evaluation_level(level)
    d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
    mf.copy_in(d_src, src)
    op.apply(d_dst, d_src, 0.0)
    tol = TOL64 if dtype == np.float64 else TOL32
    assert _rel(mf.copy_out(d_dst), ref) <= tol
    assert op.kernel_name == expect
    assert op.ghost_sides() == [0] * 12  # no neighbour is read at these levels
    # and back to the full operator
    op.set_evaluation_level(api.EVAL_ALL)
    op.apply(d_dst, d_src, 0.0)
    assert _rel(mf.copy_out(d_dst), full) <= tol
