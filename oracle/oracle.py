"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

numpy/ctypes front end of the CPU restatement of hyper.deal's advection hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.

What is restated here (reference file:line, all relative to /root/reference):

* 1-D basis data that the reference obtains from deal.II's ``ShapeInfo``
  (fe_evaluation_cell.h:93-95, fe_evaluation_cell_inverse.h:96-97,
  advection_operation.h:279-282): ``FE_DGQ(k)`` = Lagrange basis on the k+1
  Gauss-Lobatto-Legendre nodes of [0,1] (tests/tests_mf.h:162-165), ``QGauss(n_q)``
  quadrature (tests_mf.h:195-202), or GLL quadrature in collocation mode.
  deal.II (external, un-vendored, "master 2021-22") publishes these definitions;
  they are recomputed here in extended precision.
* the operator itself: ``hd_oracle.cpp`` (advection_operation.h:221-566).
* LSRK: base/time_integrators.templates.h:34-184.
* the example driver: examples/advection/include/application.h:97-560,
  source/base/time_loop.cc:34-67, operators/advection/cfl.h:58-124,
  numerics/vector_tools.h:88-220, examples/advection/cases/hyperrectangle.h.

Parity status: PINNED against examples/advection/tests/*.out (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LD = np.longdouble


# --------------------------------------------------------------------------- build
def build(force: bool = False) -> str:
    """Compile hd_oracle.cpp into oracle/_build/libhdoracle.so (g++ only)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhdoracle.so")
    src = os.path.join(_HERE, "hd_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-pthread", src, "-o", so + ".tmp"]
        try:
            subprocess.check_call(cmd)
        except subprocess.CalledProcessError:
            cmd.remove("-march=native")
            subprocess.check_call(cmd)
        os.replace(so + ".tmp", so)
    return so


def build_simd(force: bool = False) -> str:
    """Compile hd_ecl_simd.cpp (the vectorised CPU baseline of bench.py) into oracle/_build/libhdeclsimd.so."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhdeclsimd.so")
    src = os.path.join(_HERE, "hd_ecl_simd.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fno-trapping-math", src, "-o", so + ".tmp"]
        try:
            subprocess.check_call(cmd)
        except subprocess.CalledProcessError:
            cmd.remove("-march=native")
            subprocess.check_call(cmd)
        os.replace(so + ".tmp", so)
    return so


class _Mesh(ctypes.Structure):
    _fields_ = [
        ("dim_x", ctypes.c_int),
        ("dim_v", ctypes.c_int),
        ("n_cells", ctypes.c_int * 6),
        ("left", ctypes.c_double * 6),
        ("right", ctypes.c_double * 6),
        ("periodic", ctypes.c_int * 6),
    ]


_DP = ctypes.POINTER(ctypes.c_double)


class _Op(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int),
        ("nq", ctypes.c_int),
        ("S", _DP),
        ("D", _DP),
        ("Sinv", _DP),
        ("w", _DP),
        ("xq", _DP),
        ("face0", _DP),
        ("face1", _DP),
        ("skew", ctypes.c_double),
        ("velocity_kind", ctypes.c_int),
        ("a_const", _DP),
        ("a_x_table", _DP),
        ("a_v_table", _DP),
        ("bc_kind", ctypes.c_int),
        ("fn_id", ctypes.c_int),
        ("eval_level", ctypes.c_int),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.hdo_apply.argtypes = [
            ctypes.POINTER(_Mesh),
            ctypes.POINTER(_Op),
            _DP,
            _DP,
            ctypes.c_double,
            ctypes.c_int,
            ctypes.c_int64,
            ctypes.c_int64,
        ]
        _LIB.hdo_apply.restype = None
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(_DP)


_SIMD = None


class FastECL:
    """The vectorised CPU restatement of the reference's ECL kernel (oracle/hd_ecl_simd.cpp) for the benchmark
    configuration: 3D3V, degree 3, n_q = 4, periodic Cartesian lattice, constant velocity.  Same algorithm and same 1-D
    data as ``Oracle`` (which pins it in tests/test_ecl_simd_cpu.py); used by bench.py as the CPU baseline."""

    def __init__(self, n_cells, left, right, velocity, skew=0.0, nthreads=1, pin=True):
        global _SIMD
        if _SIMD is None:
            _SIMD = ctypes.CDLL(build_simd())
            ip = ctypes.POINTER(ctypes.c_int)
            _SIMD.hde_apply.argtypes = [ip, _DP, _DP, _DP, ctypes.c_double, _DP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, ctypes.c_int, ctypes.c_int]
            _SIMD.hde_apply.restype = ctypes.c_int
        assert len(n_cells) == 6
        self.n_cells = tuple(int(c) for c in n_cells)
        self.b = basis_1d(3, None, False)
        self.left = np.ascontiguousarray(left, dtype=np.float64)
        self.right = np.ascontiguousarray(right, dtype=np.float64)
        self.velocity = np.ascontiguousarray(velocity, dtype=np.float64)
        self.skew, self.nthreads, self.pin = float(skew), int(nthreads), bool(pin)
        self.ndofs = 4096 * int(np.prod(self.n_cells))

    def apply(self, src, dst=None):
        src = np.ascontiguousarray(src, dtype=np.float64)
        assert src.size == self.ndofs
        if dst is None:
            dst = np.empty_like(src)
        nc = (ctypes.c_int * 6)(*self.n_cells)
        b = self.b
        rc = _SIMD.hde_apply(nc, _ptr(self.left), _ptr(self.right), _ptr(self.velocity), self.skew, _ptr(b.S), _ptr(b.D), _ptr(b.Sinv), _ptr(b.w), _ptr(b.face0),
                             _ptr(b.face1), _ptr(src), _ptr(dst), self.nthreads, 1 if self.pin else 0)
        assert rc == 0
        return dst


# --------------------------------------------------------------------------- 1-D basis
def _legendre(n, x):
    """P_n(x), P_n'(x) in long double by the three-term recurrence."""
    x = np.asarray(x, dtype=LD)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / LD(k)
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = n * (x * p1 - p0) / (x * x - 1)  # not used at |x| = 1
    return p1, dp


def gauss_legendre(nq):
    """QGauss(nq) on [0,1]."""
    i = np.arange(nq, dtype=LD)
    x = -np.cos(LD(math.pi) * (i + LD(0.75)) / (nq + LD(0.5)))
    for _ in range(100):
        p, dp = _legendre(nq, x)
        dx = p / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    _, dp = _legendre(nq, x)
    w = 2 / ((1 - x * x) * dp * dp)
    return (x + 1) / 2, w / 2


def gauss_lobatto(n):
    """QGaussLobatto(n) on [0,1]; nodes of FE_DGQ(n-1)."""
    k = n - 1
    if n == 2:
        x = np.array([-1, 1], dtype=LD)
    else:
        xi = np.polynomial.legendre.Legendre.basis(k).deriv().roots().astype(LD)
        for _ in range(100):
            p, dp = _legendre(k, xi)
            ddp = (2 * xi * dp - k * (k + 1) * p) / (1 - xi * xi)
            dx = dp / ddp
            xi = xi - dx
            if np.max(np.abs(dx)) < 1e-19:
                break
        x = np.concatenate([[LD(-1)], np.sort(xi), [LD(1)]])
    p, _ = _legendre(k, x)
    w = 2 / (k * (k + 1) * p * p)
    return (x + 1) / 2, w / 2


def lagrange_eval(nodes, x):
    """L[q,i] = l_i(x_q) for the Lagrange basis on `nodes` (long double)."""
    nodes = np.asarray(nodes, dtype=LD)
    x = np.atleast_1d(np.asarray(x, dtype=LD))
    n = len(nodes)
    L = np.ones((len(x), n), dtype=LD)
    for i in range(n):
        for m in range(n):
            if m != i:
                L[:, i] *= (x - nodes[m]) / (nodes[i] - nodes[m])
    return L


def lagrange_deriv(nodes, x):
    """G[q,i] = l_i'(x_q)."""
    nodes = np.asarray(nodes, dtype=LD)
    x = np.atleast_1d(np.asarray(x, dtype=LD))
    n = len(nodes)
    G = np.zeros((len(x), n), dtype=LD)
    for i in range(n):
        for m in range(n):
            if m == i:
                continue
            term = np.ones(len(x), dtype=LD) / (nodes[i] - nodes[m])
            for l in range(n):
                if l != i and l != m:
                    term *= (x - nodes[l]) / (nodes[i] - nodes[l])
            G[:, i] += term
    return G


def _inv_ld(A):
    """Gauss-Jordan inverse in long double (numpy.linalg has no long double)."""
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    M = np.concatenate([A, np.eye(n, dtype=LD)], axis=1)
    for c in range(n):
        p = c + int(np.argmax(np.abs(M[c:, c])))
        M[[c, p]] = M[[p, c]]
        M[c] /= M[c, c]
        for r in range(n):
            if r != c:
                M[r] -= M[r, c] * M[c]
    return M[:, n:]


@dataclass
class Basis1D:
    n: int
    nq: int
    collocation: bool
    nodes: np.ndarray  # GLL nodes (support points of FE_DGQ)
    xq: np.ndarray
    w: np.ndarray
    S: np.ndarray  # nq x n
    D: np.ndarray  # nq x nq
    Sinv: np.ndarray  # n x nq
    face0: np.ndarray
    face1: np.ndarray
    G: np.ndarray  # nq x n   l_i'(x_q)    (used by the algebraic cross-check only)


def basis_1d(degree: int, nq: int | None = None, collocation: bool = False) -> Basis1D:
    n = degree + 1
    if nq is None:
        nq = n
    nodes, _ = gauss_lobatto(n)
    if collocation:
        assert nq == n
        xq, w = gauss_lobatto(nq)  # application.h: QGaussLobatto if DoCollocation
    else:
        xq, w = gauss_legendre(nq)
    S = lagrange_eval(nodes, xq)
    D = lagrange_deriv(xq, xq)
    f0 = lagrange_eval(xq, [0.0])[0]
    f1 = lagrange_eval(xq, [1.0])[0]
    if nq == n:
        Sinv = _inv_ld(S)
    else:
        # non-square: W-weighted L2 projection (S^T W S)^-1 S^T W  — see DESIGN.md (n_q != n)
        W = np.diag(w)
        Sinv = _inv_ld(S.T @ W @ S) @ S.T @ W
    G = lagrange_deriv(nodes, xq)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return Basis1D(n, nq, collocation, f(nodes), f(xq), f(w), f(S), f(D), f(Sinv), f(f0), f(f1), f(G))


# --------------------------------------------------------------------------- mesh / operator
@dataclass
class Mesh:
    dim_x: int
    dim_v: int
    n_cells: tuple
    left: tuple
    right: tuple
    periodic: tuple  # per direction

    @property
    def dim(self):
        return self.dim_x + self.dim_v

    @property
    def h(self):
        return [(self.right[d] - self.left[d]) / self.n_cells[d] for d in range(self.dim)]

    @property
    def total_cells(self):
        return int(np.prod(self.n_cells))

    def c_struct(self):
        m = _Mesh()
        m.dim_x, m.dim_v = self.dim_x, self.dim_v
        for d in range(6):
            m.n_cells[d] = self.n_cells[d] if d < self.dim else 1
            m.left[d] = self.left[d] if d < self.dim else 0.0
            m.right[d] = self.right[d] if d < self.dim else 1.0
            m.periodic[d] = int(self.periodic[d]) if d < self.dim else 1
        return m


def hyperrectangle_exact(points, t):
    """examples/advection/cases/hyperrectangle.h:46-57; points[..., dim]."""
    dim = points.shape[-1]
    adv = np.array([1.0, 0.15, -0.05, 0.0, 0.0, 0.0])[:dim]
    pos = points - t * adv
    r = np.sin(2.0 * pos[..., 0] * math.pi)
    for d in range(1, dim):
        r = r * np.cos(2.0 * pos[..., d] * math.pi)
    return r


HYPERRECTANGLE_VELOCITY = (1.0, 0.15, -0.05, 0.0, 0.0, 0.0)


class Oracle:
    """CPU restatement of MatrixFree + AdvectionOperation + VectorTools on a Cartesian mesh."""

    def __init__(self, mesh: Mesh, degree: int, nq: int | None = None, collocation=False, skew=0.0, velocity=None, a_x_table=None, a_v_table=None, bc_kind=1, nthreads=1,
                 eval_level=0):
        self.mesh = mesh
        self.degree = degree
        self.b = basis_1d(degree, nq, collocation)
        self.n, self.nq = self.b.n, self.b.nq
        self.skew = float(skew)
        self.nthreads = nthreads
        self.bc_kind = bc_kind
        dim = mesh.dim
        self.ndofs_cell = self.n**dim
        self.ndofs = self.ndofs_cell * mesh.total_cells
        self._keep = []
        op = _Op()
        op.n, op.nq = self.n, self.nq
        for name in ("S", "D", "Sinv", "w", "xq", "face0", "face1"):
            setattr(op, name, _ptr(getattr(self.b, name)))
        op.skew = self.skew
        if a_x_table is not None or a_v_table is not None:
            op.velocity_kind = 1
            self.a_x_table = np.ascontiguousarray(a_x_table, dtype=np.float64)
            self.a_v_table = np.ascontiguousarray(a_v_table, dtype=np.float64)
            op.a_x_table, op.a_v_table = _ptr(self.a_x_table), _ptr(self.a_v_table)
            self.velocity = None
        else:
            op.velocity_kind = 0
            self.velocity = np.ascontiguousarray(np.asarray(velocity if velocity is not None else HYPERRECTANGLE_VELOCITY[:dim], dtype=np.float64))
            assert len(self.velocity) == dim
            op.a_const = _ptr(self.velocity)
        op.bc_kind = bc_kind
        op.fn_id = 0
        op.eval_level = int(eval_level)  # AdvectionOperationEvaluationLevel: 0 all, 1 cell, 2 all_without_neighbor_load
        self._op = op
        self._mesh = mesh.c_struct()

    # -- AdvectionOperation::apply (advection_operation.h:137)
    def apply(self, src, time=0.0, cell_begin=0, cell_end=-1, dst=None):
        src = np.ascontiguousarray(src, dtype=np.float64)
        assert src.size == self.ndofs
        if dst is None:
            dst = np.zeros_like(src)
        _lib().hdo_apply(ctypes.byref(self._mesh), ctypes.byref(self._op), _ptr(src), _ptr(dst), float(time), int(self.nthreads), int(cell_begin), int(cell_end))
        return dst

    # -- helpers on the [cells..., dofs...] view
    def _view(self, vec):
        dim = self.mesh.dim
        shape = tuple(reversed(self.mesh.n_cells[:dim])) + (self.n,) * dim
        return vec.reshape(shape)

    def _points(self, ref_1d):
        """physical coordinates of tensor points: array [c_{dim-1},...,c_0, i_{dim-1},...,i_0, dim]."""
        dim = self.mesh.dim
        m = len(ref_1d)
        h = self.mesh.h
        shape = tuple(reversed(self.mesh.n_cells[:dim])) + (m,) * dim
        pts = np.zeros(shape + (dim,))
        for d in range(dim):
            x = self.mesh.left[d] + h[d] * (np.arange(self.mesh.n_cells[d])[:, None] + np.asarray(ref_1d)[None, :])  # [c, i]
            sh = [1] * (2 * dim)
            sh[dim - 1 - d] = self.mesh.n_cells[d]
            sh[2 * dim - 1 - d] = m
            pts[..., d] = x.reshape(sh)
        return pts

    # -- VectorTools::interpolate (numerics/vector_tools.h:88-137): nodal values at GLL points
    def interpolate(self, fn=hyperrectangle_exact, t=0.0):
        return np.ascontiguousarray(fn(self._points(self.b.nodes), t)).reshape(-1)

    # -- VectorTools::norm_and_error (numerics/vector_tools.h:151-220)
    def norm_and_error(self, vec, fn=hyperrectangle_exact, t=0.0):
        dim = self.mesh.dim
        u = self._view(np.asarray(vec, dtype=np.float64))
        for d in range(dim):  # S sweep along node axis of direction d
            ax = 2 * dim - 1 - d
            u = np.moveaxis(np.tensordot(self.b.S, u, axes=([1], [ax])), 0, ax)
        exact = fn(self._points(self.b.xq), t)
        h = self.mesh.h
        jxw = np.ones((self.nq,) * dim)
        for d in range(dim):
            sh = [1] * dim
            sh[dim - 1 - d] = self.nq
            jxw = jxw * (h[d] * self.b.w).reshape(sh)
        nrm = np.sum(u * u * jxw)
        err = np.sum((u - exact) ** 2 * jxw)
        return math.sqrt(nrm), math.sqrt(err)

    # -- compute_critical_time_step (operators/advection/cfl.h:58-124)
    def critical_time_step(self):
        h = self.mesh.h
        out = []
        for lo, hi in ((0, self.mesh.dim_x), (self.mesh.dim_x, self.mesh.dim)):
            vmax = max(abs(self.velocity[d] / h[d]) for d in range(lo, hi))
            out.append(1.0 / vmax if vmax > 0 else math.inf)
        return min(out)


# --------------------------------------------------------------------------- LSRK
def lsrk_coefficients(kind: str):
    """base/time_integrators.templates.h:34-86 (Kennedy, Carpenter, Lewis 2000)."""
    if kind == "rk33":
        bi = [0.245170287303492, 0.184896052186740, 0.569933660509768]
        ai = [0.755726351946097, 0.386954477304099]
    elif kind == "rk45":
        bi = [1153189308089.0 / 22510343858157.0, 1772645290293.0 / 4653164025191.0, -1672844663538.0 / 4480602732383.0, 2114624349019.0 / 3568978502595.0, 5198255086312.0 / 14908931495163.0]
        ai = [970286171893.0 / 4311952581923.0, 6584761158862.0 / 12103376702013.0, 2251764453980.0 / 15575788980749.0, 26877169314380.0 / 34165994151039.0]
    elif kind == "rk47":
        bi = [0.0941840925477795334, 0.149683694803496998, 0.285204742060440058, -0.122201846148053668, 0.0605151571191401122, 0.345986987898399296, 0.186627171718797670]
        ai = [0.241566650129646868 + bi[0], 0.0423866513027719953 + bi[1], 0.215602732678803776 + bi[2], 0.232328007537583987 + bi[3], 0.256223412574146438 + bi[4], 0.0978694102142697230 + bi[5]]
    elif kind == "rk59":
        bi = [2274579626619.0 / 23610510767302.0, 693987741272.0 / 12394497460941.0, -347131529483.0 / 15096185902911.0, 1144057200723.0 / 32081666971178.0, 1562491064753.0 / 11797114684756.0, 13113619727965.0 / 44346030145118.0, 393957816125.0 / 7825732611452.0, 720647959663.0 / 6565743875477.0, 3559252274877.0 / 14424734981077.0]
        ai = [1107026461565.0 / 5417078080134.0, 38141181049399.0 / 41724347789894.0, 493273079041.0 / 11940823631197.0, 1851571280403.0 / 6147804934346.0, 11782306865191.0 / 62590030070788.0, 9452544825720.0 / 13648368537481.0, 4435885630781.0 / 26285702406235.0, 2357909744247.0 / 11371140753790.0]
    else:
        raise NotImplementedError(kind)
    return bi, ai


def lsrk_step(op, solution, t, dt, kind="rk45"):
    """perform_time_step, base/time_integrators.templates.h:93-184; op(src, time) -> K.

    Returns the new solution (the input array is not modified)."""
    bi, ai = lsrk_coefficients(kind)
    sol = solution.copy()
    Ti = sol.copy()  # only_Ti_is_ghosted branch (:142-146)
    sum_prev_b = 0.0
    for stage in range(len(bi)):
        if stage == 0:
            c = 0.0
        else:
            c = sum_prev_b + ai[stage - 1]
            sum_prev_b += bi[stage - 1]
        K = op(Ti, t + c * dt)
        b = bi[stage] * dt
        a = 0.0 if stage == len(bi) - 1 else ai[stage] * dt
        if a == 0.0:
            sol = sol + b * K
        else:
            sol, Ti = sol + b * K, sol + a * K
    return sol


# --------------------------------------------------------------------------- example driver
def run_advection_example(json_path: str, n_points: int | None = None, nthreads: int = 1, max_lines: int | None = None):
    """Re-host of examples/advection (application.h:97-560) for the hyperrectangle case.

    Returns the list of (time, norm, error) diagnostics lines."""
    with open(json_path) as f:
        prm = json.load(f)
    g = prm["General"]
    dim_x, dim_v = int(g["DimX"]), int(g["DimV"])
    degree = int(g["DegreeX"])
    assert g.get("Case", "hyperrectangle") == "hyperrectangle"
    colloc = str(prm.get("SpatialDiscretization", {}).get("DoCollocation", "false")).lower() == "true"
    td = prm["TemporalDiscretization"]
    case = prm.get("Case", {})
    skew = float(prm.get("AdvectionOperation", {}).get("SkewFactor", 0.0))
    dim = dim_x + dim_v
    keys = ["X", "Y", "Z"]
    ncx = [int(case.get("NSubdivisionsX", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsX", 0)) for d in range(dim_x)]
    ncv = [int(case.get("NSubdivisionsV", {}).get(keys[d], 4)) * 2 ** int(case.get("NRefinementsV", 0)) for d in range(dim_v)]
    per_x = str(case.get("PeriodicX", "true")).lower() == "true"
    per_v = str(case.get("PeriodicV", "true")).lower() == "true"
    mesh = Mesh(dim_x, dim_v, tuple(ncx + ncv), (-1.0,) * dim, (1.0,) * dim, (per_x,) * dim_x + (per_v,) * dim_v)
    nq = n_points if n_points is not None else degree + 1
    orc = Oracle(mesh, degree, nq=nq, collocation=colloc, skew=skew, velocity=HYPERRECTANGLE_VELOCITY[:dim], bc_kind=1, nthreads=nthreads)

    t0, T = float(td["StartTime"]), float(td["FinalTime"])
    dt = min(float(td["TimeStep"]), float(td["CFLNumber"]) * orc.critical_time_step() / degree**1.5)  # application.h:381-384
    dt = (T - t0) / math.ceil((T - t0) / dt)  # :387-392
    tick = float(td.get("DiagnosticsTick", 0.1))
    kind = td.get("RKType", "rk45")
    max_steps = int(td.get("MaxTimeStepNumber", 10**8))

    sol = orc.interpolate(hyperrectangle_exact, 0.0)  # GLL points, application.h:326-337
    lines = []

    def diagnostics(t):
        if t != t0 and int((t + 1e-11 - t0) / tick) == int((t + 1e-11 - t0 - dt) / tick):  # :465-473
            return
        n_, e_ = orc.norm_and_error(sol, hyperrectangle_exact, t)
        lines.append((t, n_, e_))

    diagnostics(t0)
    time, step = t0 + dt, 1
    while time <= T * 1.0000000000001 and step <= max_steps:  # time_loop.cc:54-64
        sol = lsrk_step(lambda v, tt: orc.apply(v, tt), sol, time - dt, dt, kind)
        diagnostics(time)
        time += dt
        step += 1
        if max_lines is not None and len(lines) >= max_lines:
            break
    return lines


def format_line(t, nrm, err):
    """application.h:486-492."""
    return "   Time:%10.3e, norm: %17.10e, error: %17.10e" % (t, nrm, err)


def parse_golden(path):
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line.startswith("Time:"):
                continue
            parts = line.replace(",", " ").split()
            out.append((float(parts[0].split(":")[1]) if parts[0] != "Time:" else float(parts[1]), float(parts[-3]), float(parts[-1])))
    return out
