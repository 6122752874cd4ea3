"""Algebra behind the next kernel: with a velocity that varies over the TRANSVERSE quadrature
points (Vlasov-Poisson: a_x = v(q_v), a_v = grad phi(cell_x, q_x); velocity_field_view.h:111-175) the collapsed operator of
direction d is no longer one (k+1)x(k+1) matrix but

    M_a (x) C_a  +  M_|a| (x) C_abs          plus the liftings   M_a (x) L_a,f  +  M_|a| (x) L_abs,f   of BOTH neighbours,

where M_g = Sinv diag(g(q)) S acts on the transverse node indices the coefficient depends on (one v-direction for an
x-direction, all x-directions for a v-direction) and C_a, C_abs, L_a, L_abs are the speed-independent parts of the constant-
velocity matrices of basis.hpp (C = a C_a + |a| C_abs, L_f = a L_a,f + |a| L_abs,f).  This test builds that form in numpy and
compares it with the literal ECL kernel of the oracle (hd_oracle.cpp with its separable velocity tables) to round-off."""
import numpy as np
import pytest

from oracle import oracle as O


def _line_parts(b, h, s):
    """speed-independent parts of basis.hpp::direction_matrices: C = a*Ca + |a|*Cabs, L_f = a*La[f] + |a|*Labs[f]"""
    n, W = b.n, np.diag(b.w)
    B = -s * b.D + (1.0 - s) * np.linalg.inv(W) @ b.D.T @ W
    Vm = b.Sinv @ B @ b.S
    l = [b.Sinv @ (b.face0 / b.w), b.Sinv @ (b.face1 / b.w)]
    Ca, Cabs = Vm / h, np.zeros((n, n))
    La, Labs = [], []
    for f in range(2):
        nf = 1.0 if f else -1.0
        e = n - 1 if f else 0
        Ca[:, e] += -(nf / 2.0 - s * nf) / h * l[f]
        Cabs[:, e] += -1.0 / (2.0 * h) * l[f]
        La.append(-nf / (2.0 * h) * l[f])
        Labs.append(1.0 / (2.0 * h) * l[f])
    return Ca, Cabs, La, Labs


def _collapsed_apply(orc, a_x, a_v, f):
    """dense numpy evaluation of the collapsed form on a periodic mesh; f in the oracle's layout"""
    m, b, n, nq = orc.mesh, orc.b, orc.n, orc.nq
    dx, dv, dim = m.dim_x, m.dim_v, m.dim
    nc, h = m.n_cells, m.h
    u = orc._view(np.asarray(f, dtype=np.float64))  # [c_{dim-1}..c_0, i_{dim-1}..i_0]
    out = np.zeros_like(u)
    ncx, ncv = int(np.prod(nc[:dx])), int(np.prod(nc[dx:]))

    def kron_S(mats):
        r = np.array([[1.0]])
        for mm in mats:
            r = np.kron(mm, r)
        return r

    for d in range(dim):
        Ca, Cabs, La, Labs = _line_parts(b, h[d], orc.skew)
        cell_ax, node_ax = dim - 1 - d, 2 * dim - 1 - d
        for kind, Cm, Lm in (("a", Ca, La), ("abs", Cabs, Labs)):
            # line operator along d (own cell + both neighbours' end nodes)
            t = np.moveaxis(np.tensordot(Cm, u, axes=([1], [node_ax])), 0, node_ax)
            for face in range(2):
                nb = np.roll(u, 1 if face == 0 else -1, axis=cell_ax)  # lower / upper neighbour cell (periodic)
                trace = np.take(nb, n - 1 if face == 0 else 0, axis=node_ax)  # its end node facing us
                lift = np.moveaxis(np.multiply.outer(Lm[face], trace), 0, node_ax)
                t = t + lift
            # transverse coefficient matrix M_g = Sinv diag(g) S
            if d < dx:
                # g = a_x[v-cell, q_v, d] depends on the v-direction e = dx + d only (v-coordinate of the quadrature point)
                e = dx + d
                ce_ax, ne_ax = dim - 1 - e, 2 * dim - 1 - e
                res = np.zeros_like(t)
                for ce in range(nc[e]):
                    # any v-cell with coordinate ce in direction e, any q_v with the right 1-D index: take the others as 0
                    cv = ce * int(np.prod(nc[dx:e]))
                    g = np.array([a_x[cv, q * nq ** (e - dx), d] for q in range(nq)])
                    if kind == "abs":
                        g = np.abs(g)
                    M = b.Sinv @ np.diag(g) @ b.S
                    sl = [slice(None)] * (2 * dim)
                    sl[ce_ax] = ce
                    blk = t[tuple(sl)]  # cell axis e removed: node axis e shifts down by one
                    blk = np.moveaxis(np.tensordot(M, blk, axes=([1], [ne_ax - 1])), 0, ne_ax - 1)
                    res[tuple(sl)] = blk
                out += res
            else:
                # g = a_v[x-cell, q_x, d - dx] depends on all x quadrature indices: M acts on all x-node indices at once
                Sx, Six = kron_S([b.S] * dx), kron_S([b.Sinv] * dx)
                tt = t.reshape((ncv, ncx) + (n,) * dv + (n**dx,))  # [v-cells, x-cells, v-nodes.., x-nodes]
                res = np.zeros_like(tt)
                for cx in range(ncx):
                    g = a_v[cx, :, d - dx]
                    if kind == "abs":
                        g = np.abs(g)
                    M = Six @ np.diag(g) @ Sx
                    res[:, cx] = np.tensordot(tt[:, cx], M, axes=([-1], [1]))
                out += res.reshape(t.shape)
    return out.reshape(-1)


def _tables(orc, rng, mixed_signs=True):
    m, nq = orc.mesh, orc.nq
    dx, dv = m.dim_x, m.dim_v
    ncx, ncv = int(np.prod(m.n_cells[:dx])), int(np.prod(m.n_cells[dx:]))
    # a_x[v-cell, q_v, d] = coordinate of the quadrature point in v-direction d (sign changes inside the cells around v = 0)
    a_x = np.zeros((ncv, nq**dv, dx))
    for cv in range(ncv):
        for qv in range(nq**dv):
            c, q = cv, qv
            for d in range(dv):
                cd, qd = c % m.n_cells[dx + d], q % nq
                c //= m.n_cells[dx + d]
                q //= nq
                if d < dx:
                    a_x[cv, qv, d] = m.left[dx + d] + m.h[dx + d] * (cd + orc.b.xq[qd])
    a_v = rng.standard_normal((ncx, nq**dx, dv)) if mixed_signs else np.abs(rng.standard_normal((ncx, nq**dx, dv)))
    return a_x, a_v


@pytest.mark.parametrize("dx,dv,nc,skew,nq", [(1, 1, (3, 4), 0.0, None), (1, 1, (2, 3), 0.5, None), (2, 2, (2, 3, 2, 2), 0.0, None), (2, 2, (2, 2, 3, 2), 0.3, None),
                                              (1, 1, (3, 2), 0.0, 5)])
def test_collapsed_form_with_transverse_velocity_matches_literal_kernel(dx, dv, nc, skew, nq):
    dim = dx + dv
    left, right = (0.0,) * dx + (-1.3,) * dv, (2.0,) * dx + (1.7,) * dv  # v = 0 lies inside a cell
    mesh = O.Mesh(dx, dv, nc, left, right, (True,) * dim)
    rng = np.random.default_rng(11)
    probe = O.Oracle(mesh, 3, nq=nq, skew=skew, velocity=(1.0,) * dim)
    a_x, a_v = _tables(probe, rng)
    orc = O.Oracle(mesh, 3, nq=nq, skew=skew, a_x_table=a_x, a_v_table=a_v, nthreads=2)
    f = rng.standard_normal(orc.ndofs)
    ref = orc.apply(f)
    got = _collapsed_apply(orc, a_x, a_v, f)
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_collapsed_form_reduces_to_the_constant_velocity_matrices():
    """with a constant velocity M_a = a I, M_|a| = |a| I: the form is the one the shipped kernels use (basis.hpp)"""
    mesh = O.Mesh(1, 1, (3, 2), (0.0, 0.0), (1.0, 1.0), (True, True))
    vel = (0.7, -0.4)
    orc = O.Oracle(mesh, 3, skew=0.5, velocity=vel, nthreads=1)
    nq = orc.nq
    a_x = np.full((2, nq, 1), vel[0])
    a_v = np.full((3, nq, 1), vel[1])
    f = np.random.default_rng(2).standard_normal(orc.ndofs)
    assert np.max(np.abs(_collapsed_apply(orc, a_x, a_v, f) - orc.apply(f))) <= 1e-12 * np.max(np.abs(f))
