"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement of the Vlasov–Poisson right-hand side of hyper.deal (examples/vlasov_poisson) on Cartesian periodic
meshes, for the next hot-path rows of SURVEY.md §8f (general velocity field, velocity-space integration, field solve).
The advection operator itself is the literal ECL kernel of hd_oracle.cpp with its separable velocity tables
(`a_x(v-cell, q_v)`, `a_v(x-cell, q_x)`); this module adds, following the reference file:line given at each function,

* VectorTools::velocity_space_integration           numerics/vector_tools.h:238-315
* the right-hand side of the Poisson problem        examples/vlasov_poisson/include/application.h:516-565
* LaplaceOperator (symmetric interior penalty DG)   examples/vlasov_poisson/include/poisson.h:166-310
* DerivativeContainer::update (grad phi at q-points) examples/vlasov_poisson/include/derivative_container.h:157-190
* PhaseSpaceVelocityFieldView                       examples/vlasov_poisson/include/velocity_field_view.h:111-175
* compute_electric_energy, phase_space_diagnostics  examples/vlasov_poisson/include/diagnostics.h:34-143
* the driver (set-up, time loop, output line)       examples/vlasov_poisson/include/application.h:405-660,
                                                    examples/vlasov_poisson/cases/hyperrectangle.h:29-190

The reference solves the singular periodic Poisson problem by CG with a relative residual reduction of 1e-7
(poisson.h:596-600); here it is solved exactly (dense least squares on the small x-mesh), so field quantities agree with
the reference's golden output only to about that tolerance, the phase-space quantities far better.

Parity status: PINNED against examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out (tests/test_oracle_vp.py).
"""
from __future__ import annotations

import json
import math

import numpy as np

from . import oracle as O


def _kron_all(mats):
    """Kronecker product with mats[0] acting on the FASTEST index (x_0)."""
    out = np.array([[1.0]])
    for m in mats:
        out = np.kron(m, out)
    return out


class PoissonDG:
    """SIP-DG Laplacian on a periodic Cartesian mesh in the CELL-MAJOR layout of the x-space DoF vector (cells lexicographic
    with x_0 fastest, (k+1)^dim_x nodal values per cell with x_0 fastest).  poisson.h:166-250: cell term (grad u, grad v),
    interior faces  -{d_n u}[v] - [u]{d_n v} + sigma [u][v]  with sigma = (1/h_- + 1/h_+) k (k+1)  (:226-231).
    All integrands are integrated exactly by the (k+1)-point Gauss rule on Cartesian cells, so the operator is the
    Kronecker sum of 1-D matrices."""

    def __init__(self, basis: O.Basis1D, n_cells, h):
        self.b, self.n_cells, self.h = basis, tuple(n_cells), tuple(h)
        n, dim = basis.n, len(n_cells)
        self.n, self.dim = n, dim
        K1, M1 = [], []
        for d in range(dim):
            k1, m1 = self._one_d(n_cells[d], h[d])
            K1.append(k1)
            M1.append(m1)
        # line-major matrices (index = node + n * cell per direction, x_0 fastest)
        N = int(np.prod([n * c for c in n_cells]))
        K = np.zeros((N, N))
        for d in range(dim):
            K += _kron_all([K1[e] if e == d else M1[e] for e in range(dim)])
        M = _kron_all(M1)
        # permutation line-major -> cell-major
        idx = np.arange(N).reshape(tuple(reversed([n * c for c in n_cells])))  # [.., i1 + n c1, i0 + n c0] line-major index
        shape = []
        for d in reversed(range(dim)):
            shape += [n_cells[d], n]
        idx = idx.reshape(shape)  # [c_{dim-1}, i_{dim-1}, ..., c_0, i_0]
        order = [2 * j for j in range(dim)] + [2 * j + 1 for j in range(dim)]  # cells (slow..fast) then nodes (slow..fast)
        self.perm = idx.transpose(order).reshape(-1)  # cell-major position -> line-major index
        self.K = K[np.ix_(self.perm, self.perm)]
        self.M = M[np.ix_(self.perm, self.perm)]
        self.Kpinv = np.linalg.pinv(self.K, rcond=1e-12, hermitian=True)

    def _one_d(self, nc, h):
        b, n = self.b, self.n
        W = np.diag(b.w)
        Mc = h * b.S.T @ W @ b.S
        Kc = (1.0 / h) * b.G.T @ W @ b.G
        nodes = b.nodes
        v0, v1 = O.lagrange_eval(nodes, [0.0])[0], O.lagrange_eval(nodes, [1.0])[0]  # nodal basis at the two ends
        g0, g1 = O.lagrange_deriv(nodes, [0.0])[0] / h, O.lagrange_deriv(nodes, [1.0])[0] / h
        sigma = (2.0 / h) * max(self.n - 1, 1) * self.n  # (1/h + 1/h) k (k + 1)
        N = n * nc
        K, M = np.zeros((N, N)), np.zeros((N, N))
        for c in range(nc):
            s = slice(n * c, n * c + n)
            K[s, s] += Kc
            M[s, s] += Mc
        for c in range(nc):  # face between cell L = c (its xi = 1 end) and R = c + 1 (periodic), normal from L to R
            L, R = c, (c + 1) % nc
            jump = np.zeros(N)  # [u] = u_L - u_R
            jump[n * L : n * L + n] += v1
            jump[n * R : n * R + n] -= v0
            avg = np.zeros(N)  # {d_n u}
            avg[n * L : n * L + n] += 0.5 * g1
            avg[n * R : n * R + n] += 0.5 * g0
            K += -np.outer(jump, avg) - np.outer(avg, jump) + sigma * np.outer(jump, jump)
        return K, M

    def solve(self, rhs):
        """K phi = rhs on the complement of the constants (LaplaceOperator::vmult removes the mean of its argument,
        poisson.h:101-113; the potential is only used through its gradient)."""
        return self.Kpinv @ rhs


class VlasovPoissonOracle:
    """One Vlasov-Poisson right-hand side (application.h:516-600) and its diagnostics on [x-mesh] x [v-mesh]."""

    def __init__(self, dim_x, dim_v, degree, n_cells, left, right, n_points=None, nthreads=4):
        assert dim_x == dim_v, "a_x = v needs dim_x == dim_v (velocity_field_view.h:121-131)"
        self.dim_x, self.dim_v, self.dim = dim_x, dim_v, dim_x + dim_v
        self.mesh = O.Mesh(dim_x, dim_v, tuple(n_cells), tuple(left), tuple(right), (True,) * (dim_x + dim_v))
        self.h = self.mesh.h
        b = O.basis_1d(degree, n_points)
        self.b, self.n, self.nq = b, b.n, b.nq
        self.ncx, self.ncv = tuple(n_cells[:dim_x]), tuple(n_cells[dim_x:])
        self.n_cells_x, self.n_cells_v = int(np.prod(self.ncx)), int(np.prod(self.ncv))
        self.ndx, self.ndv = self.n**dim_x, self.n**dim_v
        self.nqx, self.nqv = self.nq**dim_x, self.nq**dim_v
        # a_x(v-cell, q_v) = quadrature point in v-space (velocity_field_view.h:121-131)
        a_x = np.zeros((self.n_cells_v, self.nqv, dim_x))
        for cv in range(self.n_cells_v):
            for qv in range(self.nqv):
                c, q = cv, qv
                for d in range(dim_v):
                    cd, qd = c % self.ncv[d], q % self.nq
                    c //= self.ncv[d]
                    q //= self.nq
                    a_x[cv, qv, d] = left[dim_x + d] + self.h[dim_x + d] * (cd + b.xq[qd])
        self.v_at_q = a_x.copy()
        a_v = np.zeros((self.n_cells_x, self.nqx, dim_v))
        self.adv = O.Oracle(self.mesh, degree, nq=n_points, skew=0.0, a_x_table=a_x, a_v_table=a_v, nthreads=nthreads)
        self.poisson = PoissonDG(b, self.ncx, self.h[:dim_x])
        self.potential = np.zeros(self.n_cells_x * self.ndx)
        _, self.w_gll = O.gauss_lobatto(self.n)

    # ---- layouts
    def _f_view(self, f):
        """[v-cell, x-cell, v-node, x-node]: lid = lid_x + lid_v n_cells_x (matrix_free.templates.h:553-562), x nodes fastest"""
        return np.asarray(f).reshape(self.n_cells_v, self.n_cells_x, self.ndv, self.ndx)

    def _kron(self, mats):
        return _kron_all(mats)

    # ---- numerics/vector_tools.h:238-315 with quad_no_v = 2 (Gauss-Lobatto = the nodes): rho at the x-nodes
    def velocity_space_integration(self, f):
        jxw_v = np.array([1.0])
        for d in range(self.dim_v):
            jxw_v = np.kron(self.w_gll * self.h[self.dim_x + d], jxw_v)  # lowest v-direction fastest
        rho = np.einsum("vxjn,j->xn", self._f_view(f), jxw_v)
        return rho.reshape(-1)

    # ---- application.h:516-565: (rho - mean) tested with -phi_i, mean removed again
    def poisson_rhs(self, rho):
        rho = rho - rho.mean()
        rhs = -(self.poisson.M @ rho)
        return rhs - rhs.mean()

    # ---- derivative_container.h:157-190: grad phi at the Gauss points of every x-cell  ->  a_v table
    def gradient_at_q(self, phi):
        b = self.b
        out = np.zeros((self.n_cells_x, self.nqx, self.dim_x))
        pc = phi.reshape(self.n_cells_x, self.ndx)
        for d in range(self.dim_x):
            op = self._kron([(b.G / self.h[d]) if e == d else b.S for e in range(self.dim_x)])  # [q, node]
            out[:, :, d] = pc @ op.T
        return out

    # ---- one right-hand side: steps 1-5 of application.h:516-600
    def rhs(self, f, time=0.0):
        rho = self.velocity_space_integration(f)
        self.potential = self.poisson.solve(self.poisson_rhs(rho))
        self.adv.a_v_table[...] = self.gradient_at_q(self.potential)  # negative electric field (velocity_field_view.h:134-147)
        return self.adv.apply(f, time)

    # ---- diagnostics.h:88-143: sum_q (d_d phi)^2 JxW per x-direction
    def electric_energy(self):
        g = self.gradient_at_q(self.potential)
        jxw = np.array([1.0])
        for d in range(self.dim_x):
            jxw = np.kron(self.b.w * self.h[d], jxw)
        return [float(np.sum(g[:, :, d] ** 2 * jxw[None, :])) for d in range(self.dim_x)]

    # ---- diagnostics.h:34-86: mass, l2 norm, kinetic energy, momentum at the Gauss points
    def phase_space_diagnostics(self, f):
        Sx = self._kron([self.b.S] * self.dim_x)
        Sv = self._kron([self.b.S] * self.dim_v)
        fq = np.einsum("qj,vxjn,pn->vxqp", Sv, self._f_view(f), Sx)  # [v-cell, x-cell, q_v, q_x]
        jx, jv = np.array([1.0]), np.array([1.0])
        for d in range(self.dim_x):
            jx = np.kron(self.b.w * self.h[d], jx)
        for d in range(self.dim_v):
            jv = np.kron(self.b.w * self.h[self.dim_x + d], jv)
        wf = fq * jv[None, None, :, None] * jx[None, None, None, :]
        v = self.v_at_q  # [v-cell, q_v, dim_v]
        out = [float(wf.sum()), math.sqrt(float((wf * fq).sum())), float(np.einsum("vxqp,vq->", wf, np.sum(v * v, axis=2)))]
        for d in range(self.dim_v):
            out.append(float(np.einsum("vxqp,vq->", wf, v[:, :, d])))
        while len(out) < 6:
            out.append(0.0)
        return out


def vp_initial_condition(points, dim_x):
    """examples/vlasov_poisson/cases/hyperrectangle.h:46-60: (1 + 0.01 cos(0.5 x_0)) prod_d exp(-v_d^2 / 2) / sqrt(2 pi)"""
    r = 1.0 + 0.01 * np.cos(0.5 * points[..., 0])
    for d in range(dim_x, points.shape[-1]):
        r = r * np.exp(-0.5 * points[..., d] ** 2) / math.sqrt(2.0 * math.pi)
    return r


def run_vlasov_poisson_example(json_path: str, n_points: int | None = None, nthreads: int = 4, max_steps: int | None = None):
    """examples/vlasov_poisson driver (application.h:97-700) -> rows [time, en..., mass, l2norm, kinetic, momentum...] as
    written to time_history_diagnostic.out (:640-660)."""
    prm = json.load(open(json_path))
    g, case, td = prm["General"], prm.get("Case", {}), prm["TemporalDiscretization"]
    dx, dv, k = int(g["DimX"]), int(g["DimV"]), int(g["DegreeX"])
    keys = ["X", "Y", "Z"]
    ncx = [int(case.get("NSubdivisionsX", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsX", 0)) for d in range(dx)]
    ncv = [int(case.get("NSubdivisionsV", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsV", 0)) for d in range(dv)]
    left = (0.0,) * dx + (-6.0,) * dv  # cases/hyperrectangle.h:150-165
    right = (4.0 * math.pi,) * dx + (6.0,) * dv
    vp = VlasovPoissonOracle(dx, dv, k, ncx + ncv, left, right, n_points=n_points, nthreads=nthreads)
    f = vp.adv.interpolate(lambda p, t: vp_initial_condition(p, dx), 0.0)
    # dt: application.h:421-446 with the transport direction (1,..,1, 6,..,6) of cases/hyperrectangle.h:38-43
    u = [1.0] * dx + [6.0] * dv
    h = vp.h
    crit = min(1.0 / max(abs(u[d] / h[d]) for d in rng) for rng in (range(0, dx), range(dx, dx + dv)))
    t0, T = float(td.get("StartTime", 0.0)), float(td["FinalTime"])
    dt = min(float(td.get("TimeStep", 0.1)), float(td.get("CFLNumber", 0.3)) * crit / k**1.5)
    dt = (T - t0) / math.ceil((T - t0) / dt)
    tick = float(td.get("DiagnosticsTick", 0.1))
    rk = td.get("RKType", "rk45")
    rows = []

    def diagnostics(t):
        if t != t0 and int((t + 1e-11 - t0) / tick) == int((t + 1e-11 - t0 - dt) / tick):
            return
        rows.append([t] + vp.electric_energy() + vp.phase_space_diagnostics(f))

    diagnostics(t0)
    time, step = t0 + dt, 1
    while time <= T * 1.0000000000001 and (max_steps is None or step <= max_steps):
        f = O.lsrk_step(lambda v, tt: vp.rhs(v, tt), f, time - dt, dt, rk)
        diagnostics(time)
        time += dt
        step += 1
    return rows, vp


def parse_vp_golden(path):
    rows = []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        rows.append([float(x) for x in line.split()])
    return rows
