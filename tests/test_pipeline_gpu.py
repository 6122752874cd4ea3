"""GPU parity of the pieces around the operator: LSRK stages (fused and unfused), VectorTools,
ghost-face pack + ghosted apply (two bricks emulated on one GPU), and the reference's
examples/advection golden files driven end to end through the C ABI."""
import math
import os

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


@pytest.mark.parametrize("rk", ["rk33", "rk45", "rk47", "rk59"])
def test_lsrk_fused_step_matches_oracle(api, ctx, rk):
    dx, dv, nc, k = 2, 2, (3, 2, 2, 3), 3
    left, right = (-1.0,) * 4, (1.0,) * 4
    om = O.Mesh(dx, dv, nc, left, right, (True,) * 4)
    orc = O.Oracle(om, k, skew=0.5, velocity=VEL[:4], nthreads=4)
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right)
    op = api.AdvectionOperation(mf, VEL[:4], 0.5)
    sol0 = np.random.default_rng(1).standard_normal(mf.n_dofs)
    dt = 0.004
    ref = sol0
    for s in range(3):
        ref = O.lsrk_step(lambda v, tt: orc.apply(v, tt), ref, s * dt, dt, rk)
    sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(sol, sol0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, rk)
    assert integ.n_stages() == {"rk33": 3, "rk45": 5, "rk47": 7, "rk59": 9}[rk]
    for s in range(3):
        integ.perform_time_step(sol, s * dt, dt, op)
    assert _rel(mf.copy_out(sol), ref) <= 1e-12
    # unfused path: user-supplied op callback + hd_lsrk_stage_update (the reference's structure)
    mf.copy_in(sol, sol0)
    for s in range(3):
        integ.perform_time_step(sol, s * dt, dt, lambda src, dst, t: op.apply(dst, src, t))
    assert _rel(mf.copy_out(sol), ref) <= 1e-12


@pytest.mark.parametrize("kernel", [2, 6])
@pytest.mark.parametrize("rk", ["rk45", "rk33"])
def test_lsrk_fused_fast_kernel(api, ctx, rk, kernel):
    """3D3V k=3: the fused operator+update epilogue of the two-role (2) and the three-round (6) kernel."""
    nc = (3, 2, 2, 2, 2, 2)
    left, right = (-1.0,) * 6, (1.0,) * 6
    om = O.Mesh(3, 3, nc, left, right, (True,) * 6)
    orc = O.Oracle(om, 3, skew=0.5, velocity=VEL, nthreads=8)
    mf = api.MatrixFree(ctx, 3, 3, 3, nc, left, right)
    op = api.AdvectionOperation(mf, VEL, 0.5)
    op.set_kernel(kernel)
    sol0 = np.random.default_rng(2).standard_normal(mf.n_dofs)
    dt = 0.002
    ref = sol0
    for s in range(2):
        ref = O.lsrk_step(lambda v, tt: orc.apply(v, tt), ref, s * dt, dt, rk)
    sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(sol, sol0)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, rk)
    for s in range(2):
        integ.perform_time_step(sol, s * dt, dt, op)
    assert op.kernel_name == {2: "advect_3d3v_k3_fused_lsrk", 6: "rounds_3d3v_k3_fused_lsrk"}[kernel]
    assert _rel(mf.copy_out(sol), ref) <= 1e-12


def test_lsrk_scalar_ode(api, ctx):
    """tests/time_discretization/time_integrators_02.cc: y' = y sin^2 t, rk45, dt = .1, 100 steps -> 118.127."""
    mf = api.MatrixFree(ctx, 1, 1, 1, (1, 1), (0.0, 0.0), (1.0, 1.0))
    y, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    mf.copy_in(y, np.ones(mf.n_dofs))
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, "rk45")

    def rhs(src, dst, t):
        mf.copy_in(dst, mf.copy_out(src) * math.sin(t) ** 2)

    for it in range(100):
        integ.perform_time_step(y, 0.1 * it, 0.1, rhs)
    out = mf.copy_out(y)
    assert all("%.6g" % v == "118.127" for v in out)


@pytest.mark.parametrize("dx,dv,nc,k,nq", [(1, 1, (16, 16), 3, None), (2, 2, (4, 4, 4, 4), 3, None), (2, 2, (4, 3, 2, 4), 3, 5), (3, 3, (2, 2, 2, 2, 2, 2), 3, None)])
def test_interpolate_and_norm(api, ctx, dx, dv, nc, k, nq):
    dim = dx + dv
    left, right = (-1.0,) * dim, (1.0,) * dim
    om = O.Mesh(dx, dv, nc, left, right, (True,) * dim)
    orc = O.Oracle(om, k, nq=nq, velocity=VEL[:dim])
    mf = api.MatrixFree(ctx, dx, dv, k, nc, left, right, n_points=nq)
    v = mf.initialize_dof_vector()
    api.VectorTools.interpolate(mf, v, api.FN_HYPERRECTANGLE, 0.13)
    ref = orc.interpolate(O.hyperrectangle_exact, 0.13)
    got = mf.copy_out(v)
    assert np.max(np.abs(got - ref)) <= 1e-13
    n_ref, e_ref = orc.norm_and_error(ref, O.hyperrectangle_exact, 0.13)
    n_gpu, e_gpu = api.VectorTools.norm_and_error(mf, v, api.FN_HYPERRECTANGLE, 0.13)
    assert abs(n_gpu - n_ref) <= 1e-12 * n_ref
    assert abs(e_gpu - e_ref) <= 1e-9 * e_ref + 1e-15


def test_vector_tools_reference_output(api, ctx):
    """tests/vector_tools/vector_tools_01.mpirun=1.output: 1D1V k=3, sin*cos on [-1,1]^2 -> norm 0.5... (here: the
    convergence order of the interpolation error is 4 when refining 16 -> 32 -> 64 cells per direction)."""
    errs = []
    for n in (16, 32, 64):
        mf = api.MatrixFree(ctx, 1, 1, 3, (n, n), (-1.0, -1.0), (1.0, 1.0))
        v = mf.initialize_dof_vector()
        api.VectorTools.interpolate(mf, v, api.FN_HYPERRECTANGLE, 0.0)
        nrm, err = api.VectorTools.norm_and_error(mf, v, api.FN_HYPERRECTANGLE, 0.0)
        assert abs(nrm - 1.0) < 1e-5  # ||sin cos||_L2([-1,1]^2) = 1
        errs.append(err)
    assert 3.9 < math.log2(errs[0] / errs[1]) < 4.1 and 3.9 < math.log2(errs[1] / errs[2]) < 4.1


@pytest.mark.parametrize("split_dir,kernel,parts", [(0, 1, False), (2, 1, False), (5, 1, False), (0, 2, False), (1, 2, False), (2, 2, False), (3, 2, False), (4, 2, False), (5, 2, False),
                                                    (0, 2, True), (1, 2, True), (2, 2, True), (4, 2, True), (5, 2, True), (2, 1, True),
                                                    (0, 6, False), (1, 6, False), (2, 6, False), (3, 6, False), (4, 6, False), (5, 6, False),
                                                    (0, 6, True), (1, 6, True), (2, 6, True), (3, 6, True), (4, 6, True), (5, 6, True)])
def test_two_bricks_with_ghost_faces(api, ctx, split_dir, kernel, parts):
    """Partition the lattice into two bricks along one direction, exchange packed faces by hand and
    compare with the unpartitioned operator (the ghost path of matrix_free/vector_partitioner.h)."""
    dx, dv, k = 3, 3, 3
    nc = [2, 2, 2, 2, 2, 2]
    nc[split_dir] = 4
    left, right = (-1.0,) * 6, (1.0,) * 6
    om = O.Mesh(dx, dv, tuple(nc), left, right, (True,) * 6)
    orc = O.Oracle(om, k, skew=0.5, velocity=VEL, nthreads=8)
    src = np.random.default_rng(9).standard_normal(orc.ndofs)
    ref = orc.apply(src)
    nd = 4**6
    full = src.reshape(tuple(reversed(nc)) + (nd,))
    ref_full = ref.reshape(tuple(reversed(nc)) + (nd,))
    axis = 5 - split_dir
    bricks = []
    for b in range(2):
        loc = list(nc)
        loc[split_dir] = 2
        off = [0] * 6
        off[split_dir] = 2 * b
        side_kind = [[api.SIDE_PERIODIC_LOCAL] * 2 for _ in range(6)]
        side_kind[split_dir] = [api.SIDE_GHOST, api.SIDE_GHOST]
        mf = api.MatrixFree(ctx, dx, dv, k, loc, left, right, n_cells_global=nc, cell_offset=off, side_kind=side_kind)
        sl = [slice(None)] * 7
        sl[axis] = slice(2 * b, 2 * b + 2)
        u = np.ascontiguousarray(full[tuple(sl)]).reshape(-1)
        d_src, d_dst = mf.initialize_dof_vector(), mf.initialize_dof_vector()
        mf.copy_in(d_src, u)
        assert mf.halo_total == 2 * 32 * 4**5
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
            # my (dir, side) ghosts = the neighbour's boundary layer on its opposite side; with two
            # bricks on a periodic direction both of my neighbours are the other brick
            o_ot = other["mf"].halo_offset(split_dir, 1 - side)
            ghost[o_me : o_me + n_me] = other["send"][o_ot : o_ot + n_me]
        op = api.AdvectionOperation(mf, VEL, 0.5)
        op.set_kernel(kernel)
        needed = op.ghost_sides()
        assert sum(needed) == 1 and needed[2 * split_dir + (0 if VEL[split_dir] > 0 else 1)] == 1
        for side in range(2):
            if not needed[2 * split_dir + side]:  # the outflow side is never read: poison it
                o_me, n_me = mf.halo_offset(split_dir, side), mf.ghost_size(split_dir, side)
                ghost[o_me : o_me + n_me] = np.nan
        mf.copy_in(me["ghost"], ghost)
        if parts:
            # interior cells first (they must not touch the ghost buffer: poison it meanwhile), then the boundary layer
            mf.copy_in(me["ghost"], np.full(mf.halo_total, np.nan))
            op.apply_part(me["dst"], me["src"], 0.0, me["ghost"], api.PART_INTERIOR)
            mf.copy_in(me["ghost"], ghost)
            op.apply_part(me["dst"], me["src"], 0.0, me["ghost"], api.PART_BOUNDARY)
        else:
            op.apply(me["dst"], me["src"], 0.0, ghosts=me["ghost"])
        assert op.kernel_name == {1: "generic", 2: "advect_3d3v_k3", 6: "rounds_3d3v_k3"}[kernel]
        out = mf.copy_out(me["dst"])
        expect = np.ascontiguousarray(ref_full[me["sl"]]).reshape(-1)
        assert _rel(out, expect) <= 1e-12


def _run_example_gpu(api, ctx, json_path, n_points):
    """examples/advection driver (application.h:97-560) with every vector on the device."""
    import json

    prm = json.load(open(json_path))
    g = prm["General"]
    dx, dv, k = int(g["DimX"]), int(g["DimV"]), int(g["DegreeX"])
    dim = dx + dv
    colloc = str(prm.get("SpatialDiscretization", {}).get("DoCollocation", "false")).lower() == "true"
    td, case = prm["TemporalDiscretization"], prm.get("Case", {})
    skew = float(prm.get("AdvectionOperation", {}).get("SkewFactor", 0.0))
    keys = ["X", "Y", "Z"]
    ncx = [int(case.get("NSubdivisionsX", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsX", 0)) for d in range(dx)]
    ncv = [int(case.get("NSubdivisionsV", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsV", 0)) for d in range(dv)]
    per_x = str(case.get("PeriodicX", "true")).lower() == "true"
    per_v = str(case.get("PeriodicV", "true")).lower() == "true"
    per = (per_x,) * dx + (per_v,) * dv
    nc = ncx + ncv
    vel = np.array(O.HYPERRECTANGLE_VELOCITY[:dim])
    mf = api.MatrixFree(ctx, dx, dv, k, nc, (-1.0,) * dim, (1.0,) * dim, periodic=per, n_points=n_points, collocation=colloc)
    op = api.AdvectionOperation(mf, vel, skew)
    op.set_dirichlet_builtin(api.FN_HYPERRECTANGLE)
    h = [2.0 / c for c in nc]
    crit = min(1.0 / max(abs(vel[d] / h[d]) for d in rng) if max(abs(vel[d]) for d in rng) > 0 else math.inf for rng in (range(0, dx), range(dx, dim)))
    t0, T = float(td["StartTime"]), float(td["FinalTime"])
    dt = min(float(td["TimeStep"]), float(td["CFLNumber"]) * crit / k**1.5)
    dt = (T - t0) / math.ceil((T - t0) / dt)
    tick = float(td.get("DiagnosticsTick", 0.1))
    sol, Ki, Ti = (mf.initialize_dof_vector() for _ in range(3))
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki, Ti, td.get("RKType", "rk45"))
    api.VectorTools.interpolate(mf, sol, api.FN_HYPERRECTANGLE, 0.0)
    lines = []

    def diag(t):
        if t != t0 and int((t + 1e-11 - t0) / tick) == int((t + 1e-11 - t0 - dt) / tick):
            return
        lines.append((t,) + api.VectorTools.norm_and_error(mf, sol, api.FN_HYPERRECTANGLE, t))

    diag(t0)
    time = t0 + dt
    while time <= T * 1.0000000000001:
        integ.perform_time_step(sol, time - dt, dt, op)
        diag(time)
        time += dt
    return lines


# (2D2V _04/_06/_08, _q5..._04 and 1D1V _04 are the reference's UseECL = false / DoBuffering = true runs: one ECL-style kernel serves both loop types here)
GOLDEN = ["adv_2D_2D_k3.hyperrectangle_%02d" % i for i in (1, 3, 5, 7)] + ["adv_2D_2D_k3_q5.hyperrectangle_01", "adv_2D_2D_k3_q5.hyperrectangle_03", "adv_1D_1D_k3.hyperrectangle_01", "adv_1D_1D_k3.hyperrectangle_01_rk33", "adv_1D_1D_k3.hyperrectangle_02"] + \
         ["adv_2D_2D_k3.hyperrectangle_%02d" % i for i in (2, 4, 6, 8)] + ["adv_2D_2D_k3_q5.hyperrectangle_04", "adv_1D_1D_k3.hyperrectangle_04"]


@pytest.mark.parametrize("name", GOLDEN)
def test_reference_golden_files_on_gpu(api, ctx, golden_dir, name):
    """The reference's own end-to-end goldens (examples/advection/tests/*.out; numdiff -a 1e-5 -r 1e-8
    in the reference, asserted here at the 11 printed digits)."""
    conf = open(os.path.join(golden_dir, name.split(".")[0] + ".configuration")).read()
    nq = int(conf.split("N_POINTS=")[1].split()[0])
    lines = _run_example_gpu(api, ctx, os.path.join(golden_dir, name + ".json"), nq)
    gold = O.parse_golden(os.path.join(golden_dir, name + ".out"))
    assert len(lines) == len(gold)
    for (t1, n1, e1), (t2, n2, e2) in zip(lines, gold):
        assert abs(t1 - t2) <= 1e-3 * max(abs(t2), 1e-3)
        assert abs(n1 - n2) <= 2e-10 * abs(n2), (name, t1, n1, n2)
        if e2 > 1e-12:
            assert abs(e1 - e2) <= 2e-10 * abs(e2), (name, t1, e1, e2)
        else:
            assert e1 < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_halo_pack_selective_and_direct(api, ctx, dtype):
    """hd_halo_pack_ex: masked segments stay untouched; a segment can be written straight into another
    buffer (the peer's ghost segment on a multi-GPU box) instead of the send buffer."""
    nc = (2, 3, 2, 2)
    side_kind = [[api.SIDE_GHOST, api.SIDE_GHOST], [api.SIDE_PERIODIC_LOCAL] * 2, [api.SIDE_GHOST, api.SIDE_GHOST], [api.SIDE_PERIODIC_LOCAL] * 2]
    mf = api.MatrixFree(ctx, 2, 2, 3, nc, (0.0,) * 4, (1.0,) * 4, n_cells_global=(4, 3, 4, 2), cell_offset=(2, 0, 0, 0), side_kind=side_kind, dtype=dtype)
    src = np.random.default_rng(4).standard_normal(mf.n_dofs).astype(dtype)
    d_src, d_a, d_b, d_peer = (mf.initialize_dof_vector() for _ in range(4))
    mf.copy_in(d_src, src)
    mf.halo_pack(d_src, d_a)
    full = mf.copy_out(d_a, mf.halo_total)
    # reference layout: face cells lexicographic, nodal layer 0 / k of direction d
    u = src.reshape(tuple(reversed(nc)) + (4,) * 4)
    for d in (0, 2):
        for side in range(2):
            sl = [slice(None)] * 8
            sl[3 - d] = -1 if side else 0
            sl[4 + 3 - d] = -1 if side else 0
            o, n = mf.halo_offset(d, side), mf.ghost_size(d, side)
            assert np.array_equal(full[o : o + n], np.ascontiguousarray(u[tuple(sl)]).reshape(-1))
    mask = [0] * 12
    mask[2 * 2 + 1] = 1
    mask[2 * 0 + 0] = 1
    peer = [0] * 12
    itemsize = np.dtype(dtype).itemsize
    peer[2 * 0 + 0] = d_peer + 64 * itemsize  # direction 0 is the strided (8-byte) gather path
    mf.copy_in(d_b, np.full(mf.halo_total, 5.0, dtype=dtype))
    mf.halo_pack(d_src, d_b, send_mask=mask, peer_dst=peer)
    got = mf.copy_out(d_b, mf.halo_total)
    o, n = mf.halo_offset(2, 1), mf.ghost_size(2, 1)
    assert np.array_equal(got[o : o + n], full[o : o + n])
    untouched = np.ones(mf.halo_total, dtype=bool)
    untouched[o : o + n] = False
    assert np.all(got[untouched] == 5.0)
    o0, n0 = mf.halo_offset(0, 0), mf.ghost_size(0, 0)
    assert np.array_equal(mf.copy_out(d_peer, 64 + n0)[64:], full[o0 : o0 + n0])


@pytest.mark.parametrize("kernel", [2, 6])
@pytest.mark.parametrize("dirs", [(0,), (1,), (2,), (3,), (4,), (5,), (0, 2, 5), (1, 3, 4)])
@pytest.mark.parametrize("vel", [(1.0, 0.15, -0.05, 0.1, -0.15, 0.5), (-1.0, -0.15, 0.05, -0.1, 0.15, -0.5)])
def test_fused_halo_kernel_self_exchange(api, ctx, dirs, vel, kernel):
    """hd_advection_apply_overlapped: operator + ghost exchange in one kernel.  One brick whose periodic
    neighbour is itself: the halo warp of every CTA stores the brick's own boundary layers into its own ghost
    segments and bumps its own arrival counters, the boundary phase waits for them.  Must equal the plain
    periodic operator bit for bit."""
    import torch

    nc = (4, 3, 2, 3, 2, 3)
    left, right = (-1.0,) * 6, (1.0,) * 6
    mf0 = api.MatrixFree(ctx, 3, 3, 3, nc, left, right)
    op0 = api.AdvectionOperation(mf0, vel, 0.5)
    op0.set_kernel(kernel)
    u = np.random.default_rng(21).standard_normal(mf0.n_dofs)
    d_src, d_ref = mf0.initialize_dof_vector(), mf0.initialize_dof_vector()
    mf0.copy_in(d_src, u)
    op0.apply(d_ref, d_src, 0.0)
    ref = mf0.copy_out(d_ref)
    side_kind = [[api.SIDE_GHOST, api.SIDE_GHOST] if d in dirs else [api.SIDE_PERIODIC_LOCAL] * 2 for d in range(6)]
    mf = api.MatrixFree(ctx, 3, 3, 3, nc, left, right, side_kind=side_kind)
    op = api.AdvectionOperation(mf, vel, 0.5)
    op.set_kernel(kernel)
    needed = op.ghost_sides()
    ghost = torch.full((mf.halo_total,), float("nan"), dtype=torch.float64, device="cuda")
    counters = torch.zeros(12, dtype=torch.int32, device="cuda")
    dst = torch.zeros(mf.n_dofs, dtype=torch.float64, device="cuda")
    # my layer (d, side) is my own ghost (d, 1 - side); only the sides the operator reads are sent
    sends = [(d, s, ghost.data_ptr() + 8 * mf.halo_offset(d, 1 - s), counters.data_ptr() + 4 * (2 * d + (1 - s))) for d in dirs for s in range(2) if needed[2 * d + (1 - s)]]
    assert len(sends) == len(dirs)
    for epoch in (1, 2, 3):
        ghost.fill_(float("nan"))
        op.apply_overlapped(dst.data_ptr(), d_src, 0.0, ghost.data_ptr(), sends, counters.data_ptr(), epoch * op.n_halo_senders)
        torch.cuda.synchronize()
        assert not op.overlap_timed_out()
        assert op.kernel_name == {2: "advect_3d3v_k3", 6: "rounds_3d3v_k3"}[kernel]
        assert np.array_equal(dst.cpu().numpy(), ref)
    # configurations the fused kernel does not cover are refused, not silently mis-computed
    mf2 = api.MatrixFree(ctx, 2, 2, 3, (2, 2, 2, 2), (0.0,) * 4, (1.0,) * 4, side_kind=[[api.SIDE_GHOST] * 2] + [[api.SIDE_PERIODIC_LOCAL] * 2] * 3)
    op2 = api.AdvectionOperation(mf2, vel[:4], 0.0)
    v = mf2.initialize_dof_vector()
    with pytest.raises(api.HdError):
        op2.apply_overlapped(v, mf2.initialize_dof_vector(), 0.0, v, [], counters.data_ptr(), 1)
