"""ctypes binding of libhdgpu.so (include/hyperdeal_b200.h) for the Python-side plumbing:
tests, bench.py and the torch.distributed launcher.  The product host layer is the C++ shim in
hyperdeal_b200/cpp/; this module only forwards to the same C ABI and never computes anything
itself.  If the library or a CUDA device is missing every call raises — there is no CPU
fallback (the CPU oracle lives in oracle/ and is test infrastructure only).

Class and method names follow the reference:
  MatrixFree            hyperdeal::MatrixFree             (matrix_free/matrix_free.h:39)
  AdvectionOperation    hyperdeal::advection::AdvectionOperation (operators/advection/advection_operation.h:56)
  LowStorageRungeKuttaIntegrator                           (base/time_integrators.h:48)
  VectorTools.interpolate / norm_and_error                 (numerics/vector_tools.h:88, :151)
"""
from __future__ import annotations

import ctypes
import math
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_void_p

import numpy as np

HD_MAX_DIM = 6
HD_F64, HD_F32 = 0, 1
SIDE_PERIODIC_LOCAL, SIDE_GHOST, SIDE_DIRICHLET, SIDE_DIRICHLET_HOM = 0, 1, 2, 3
FN_ZERO, FN_HYPERRECTANGLE = 0, 1
EVAL_ALL, EVAL_CELL, EVAL_ALL_WITHOUT_NEIGHBOR_LOAD = 0, 1, 2
PART_ALL, PART_INTERIOR, PART_BOUNDARY = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HD_LIBHDGPU", os.path.join(_HERE, "lib", "libhdgpu.so"))  # HD_LIBHDGPU: use another build of the same C ABI

EXPORTS = [
    "hd_last_error", "hd_version", "hd_context_create", "hd_context_destroy", "hd_context_set_stream",
    "hd_context_synchronize", "hd_device_count", "hd_device_malloc", "hd_device_free", "hd_mesh_create", "hd_mesh_destroy", "hd_mesh_n_dofs",
    "hd_mesh_n_cells", "hd_mesh_dofs_per_cell", "hd_mesh_ghost_size", "hd_mesh_basis", "hd_vector_alloc",
    "hd_vector_free", "hd_vector_copy", "hd_vector_copy_in", "hd_vector_copy_out", "hd_vector_zero", "hd_vector_zero_n", "hd_vector_copy_n", "hd_advection_create",
    "hd_advection_destroy", "hd_advection_set_phase_space_velocity", "hd_advection_apply", "hd_advection_apply_part", "hd_advection_apply_overlapped", "hd_advection_overlap_status", "hd_advection_n_ctas", "hd_advection_n_halo_senders", "hd_advection_set_halo_senders", "hd_stream_write_flag", "hd_stream_wait_flag", "hd_advection_ghost_sides", "hd_advection_apply_host", "hd_advection_set_kernel", "hd_advection_set_l2_hints", "hd_advection_set_row_tile",
    "hd_advection_kernel_name", "hd_advection_launch_count", "hd_advection_set_evaluation_level",
    "hd_ipc_export", "hd_ipc_open", "hd_ipc_close", "hd_advection_set_dirichlet_values",
    "hd_advection_set_dirichlet_builtin", "hd_halo_pack", "hd_halo_pack_ex", "hd_halo_offset", "hd_halo_total", "hd_lsrk_create",
    "hd_lsrk_destroy", "hd_lsrk_n_stages", "hd_lsrk_coefficients", "hd_lsrk_stage_update", "hd_lsrk_step", "hd_lsrk_stage_fused", "hd_lsrk_stage_overlapped",
    "hd_multi_create", "hd_multi_destroy", "hd_multi_n_gpus", "hd_multi_mesh", "hd_multi_context", "hd_multi_n_dofs", "hd_multi_synchronize", "hd_multi_vector_alloc",
    "hd_multi_vector_free", "hd_multi_vector_copy_in", "hd_multi_vector_copy_out", "hd_multi_interpolate_builtin", "hd_multi_norm_and_error_builtin",
    "hd_multi_advection_create", "hd_multi_advection_destroy", "hd_multi_advection_set_dirichlet_builtin", "hd_multi_advection_kernel_name", "hd_multi_advection_apply",
    "hd_multi_lsrk_create", "hd_multi_lsrk_destroy", "hd_multi_lsrk_step",
    "hd_interpolate_builtin", "hd_norm_and_error_builtin", "hd_mesh_n_dofs_x", "hd_vector_alloc_x", "hd_velocity_space_integration", "hd_poisson_create", "hd_poisson_destroy", "hd_poisson_solve", "hd_poisson_last_solve", "hd_poisson_potential", "hd_phase_space_diagnostics", "hd_field_energy", "hd_timer_start", "hd_timer_stop",
]


class MeshDesc(ctypes.Structure):
    _fields_ = [
        ("dim_x", c_int), ("dim_v", c_int), ("degree", c_int), ("n_points", c_int), ("collocation", c_int),
        ("number_type", c_int),
        ("left", c_double * HD_MAX_DIM), ("right", c_double * HD_MAX_DIM),
        ("n_cells_global", c_int * HD_MAX_DIM), ("n_cells", c_int * HD_MAX_DIM), ("cell_offset", c_int * HD_MAX_DIM),
        ("side_kind", (c_int * 2) * HD_MAX_DIM),
    ]


class HaloSend(ctypes.Structure):
    """hd_halo_send: boundary layer (dir, side) -> peer-mapped ghost segment + the peer's arrival counter."""

    _fields_ = [("dir", c_int), ("side", c_int), ("dst", c_void_p), ("arrival_counter", c_void_p)]


class HdError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libhdgpu.so (built in-tree by hyperdeal_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HdError(f"{LIB_PATH} is missing: run `python -m hyperdeal_b200.build` (needs nvcc); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    L.hd_last_error.restype = c_char_p
    L.hd_advection_kernel_name.restype = c_char_p
    L.hd_advection_kernel_name.argtypes = [c_void_p]
    for name in ("hd_mesh_n_dofs", "hd_mesh_n_dofs_x", "hd_mesh_n_cells", "hd_halo_total", "hd_advection_launch_count"):
        getattr(L, name).restype = c_int64
        getattr(L, name).argtypes = [c_void_p]
    for name in ("hd_mesh_ghost_size", "hd_halo_offset"):
        getattr(L, name).restype = c_int64
        getattr(L, name).argtypes = [c_void_p, c_int, c_int]
    L.hd_context_create.argtypes = [c_int, POINTER(c_void_p)]
    L.hd_context_destroy.argtypes = [c_void_p]
    L.hd_context_set_stream.argtypes = [c_void_p, c_void_p]
    L.hd_context_synchronize.argtypes = [c_void_p]
    L.hd_device_count.argtypes = [POINTER(c_int)]
    L.hd_mesh_create.argtypes = [c_void_p, POINTER(MeshDesc), POINTER(c_void_p)]
    L.hd_mesh_destroy.argtypes = [c_void_p]
    L.hd_mesh_dofs_per_cell.argtypes = [c_void_p]
    L.hd_mesh_basis.argtypes = [c_void_p, c_int, c_void_p]
    L.hd_vector_alloc.argtypes = [c_void_p, c_int, POINTER(c_void_p)]
    L.hd_vector_free.argtypes = [c_void_p, c_void_p]
    L.hd_vector_copy_in.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
    L.hd_vector_copy_out.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
    L.hd_vector_zero.argtypes = [c_void_p, c_void_p]
    L.hd_vector_copy.argtypes = [c_void_p, c_void_p, c_void_p]
    L.hd_advection_create.argtypes = [c_void_p, c_double, POINTER(c_double), POINTER(c_void_p)]
    L.hd_advection_destroy.argtypes = [c_void_p]
    L.hd_advection_set_phase_space_velocity.argtypes = [c_void_p, c_void_p]
    L.hd_advection_apply.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double]
    L.hd_advection_apply_part.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int]
    L.hd_advection_apply_overlapped.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, POINTER(HaloSend), c_int, c_void_p, c_int]
    L.hd_advection_overlap_status.argtypes = [c_void_p, POINTER(c_int)]
    L.hd_stream_write_flag.argtypes = [c_void_p, c_void_p, c_int]
    L.hd_device_malloc.argtypes = [c_void_p, ctypes.c_size_t, POINTER(c_void_p)]
    L.hd_device_free.argtypes = [c_void_p, c_void_p]
    L.hd_ipc_export.argtypes = [c_void_p, c_void_p, c_void_p]
    L.hd_ipc_open.argtypes = [c_void_p, c_void_p, POINTER(c_void_p)]
    L.hd_ipc_close.argtypes = [c_void_p, c_void_p]
    L.hd_stream_wait_flag.argtypes = [c_void_p, c_void_p, c_int]
    L.hd_advection_ghost_sides.argtypes = [c_void_p, POINTER(c_int)]
    L.hd_halo_pack_ex.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_int), POINTER(c_void_p), c_void_p, POINTER(c_int)]
    L.hd_advection_n_ctas.argtypes = [c_void_p]
    L.hd_advection_n_halo_senders.argtypes = [c_void_p]
    L.hd_advection_set_halo_senders.argtypes = [c_void_p, c_int]
    L.hd_advection_apply_host.argtypes = [c_void_p, c_void_p, c_void_p, c_double]
    L.hd_advection_set_kernel.argtypes = [c_void_p, c_int]
    L.hd_advection_set_evaluation_level.argtypes = [c_void_p, c_int]
    L.hd_advection_set_l2_hints.argtypes = [c_void_p, c_int]
    L.hd_advection_set_row_tile.argtypes = [c_void_p, POINTER(c_int)]
    L.hd_advection_set_dirichlet_values.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int64]
    L.hd_advection_set_dirichlet_builtin.argtypes = [c_void_p, c_int]
    L.hd_halo_pack.argtypes = [c_void_p, c_void_p, c_void_p]
    L.hd_lsrk_create.argtypes = [c_void_p, c_char_p, POINTER(c_void_p)]
    L.hd_lsrk_destroy.argtypes = [c_void_p]
    L.hd_lsrk_n_stages.argtypes = [c_void_p]
    L.hd_lsrk_coefficients.argtypes = [c_void_p, c_int, c_void_p]
    L.hd_lsrk_stage_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double]
    L.hd_lsrk_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double]
    L.hd_interpolate_builtin.argtypes = [c_void_p, c_void_p, c_int, c_double]
    L.hd_vector_alloc_x.argtypes = [c_void_p, POINTER(c_void_p)]
    L.hd_poisson_create.argtypes = [c_void_p, POINTER(c_void_p)]
    L.hd_poisson_destroy.argtypes = [c_void_p]
    L.hd_poisson_solve.argtypes = [c_void_p, c_void_p, c_void_p, c_double, c_int, POINTER(c_int)]
    L.hd_poisson_last_solve.argtypes = [c_void_p, POINTER(c_int), POINTER(c_double)]
    L.hd_vector_zero_n.argtypes = [c_void_p, c_void_p, c_int64]
    L.hd_vector_copy_n.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
    L.hd_lsrk_stage_fused.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double]
    L.hd_lsrk_stage_overlapped.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(HaloSend), c_int, c_void_p, c_int, c_double, c_double]
    L.hd_multi_create.argtypes = [c_int, POINTER(c_int), POINTER(MeshDesc), POINTER(c_int), POINTER(c_void_p)]
    L.hd_multi_destroy.argtypes = [c_void_p]
    L.hd_multi_n_gpus.argtypes = [c_void_p]
    L.hd_multi_mesh.argtypes = [c_void_p, c_int]
    L.hd_multi_mesh.restype = c_void_p
    L.hd_multi_context.argtypes = [c_void_p, c_int]
    L.hd_multi_context.restype = c_void_p
    L.hd_multi_n_dofs.argtypes = [c_void_p]
    L.hd_multi_n_dofs.restype = c_int64
    L.hd_multi_synchronize.argtypes = [c_void_p]
    L.hd_multi_vector_alloc.argtypes = [c_void_p, POINTER(c_void_p)]
    L.hd_multi_vector_free.argtypes = [c_void_p, POINTER(c_void_p)]
    L.hd_multi_vector_copy_in.argtypes = [c_void_p, POINTER(c_void_p), c_void_p]
    L.hd_multi_vector_copy_out.argtypes = [c_void_p, POINTER(c_void_p), c_void_p]
    L.hd_multi_interpolate_builtin.argtypes = [c_void_p, POINTER(c_void_p), c_int, c_double]
    L.hd_multi_norm_and_error_builtin.argtypes = [c_void_p, POINTER(c_void_p), c_int, c_double, POINTER(c_double)]
    L.hd_multi_advection_create.argtypes = [c_void_p, c_double, POINTER(c_double), POINTER(c_void_p)]
    L.hd_multi_advection_destroy.argtypes = [c_void_p]
    L.hd_multi_advection_set_dirichlet_builtin.argtypes = [c_void_p, c_int]
    L.hd_multi_advection_kernel_name.argtypes = [c_void_p]
    L.hd_multi_advection_kernel_name.restype = c_char_p
    L.hd_multi_advection_apply.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), c_double]
    L.hd_multi_lsrk_create.argtypes = [c_void_p, c_char_p, POINTER(c_void_p)]
    L.hd_multi_lsrk_destroy.argtypes = [c_void_p]
    L.hd_multi_lsrk_step.argtypes = [c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_double, c_double]
    L.hd_poisson_potential.argtypes = [c_void_p]
    L.hd_poisson_potential.restype = c_void_p
    L.hd_phase_space_diagnostics.argtypes = [c_void_p, c_void_p, POINTER(c_double)]
    L.hd_field_energy.argtypes = [c_void_p, c_void_p, POINTER(c_double)]
    L.hd_velocity_space_integration.argtypes = [c_void_p, c_void_p, c_void_p]
    L.hd_norm_and_error_builtin.argtypes = [c_void_p, c_void_p, c_int, c_double, POINTER(c_double)]
    L.hd_timer_start.argtypes = [c_void_p]
    L.hd_timer_stop.argtypes = [c_void_p, POINTER(c_double)]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise HdError("libhdgpu: %s (code %d)" % (lib().hd_last_error().decode(), rc))
    return rc


class Context:
    """One GPU + stream (stands in for the (comm, comm_sm) pair, matrix_free.h:108)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = c_void_p()
        _check(lib().hd_context_create(device, byref(self._h)))
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream: int):
        _check(lib().hd_context_set_stream(self._h, c_void_p(stream)))

    def synchronize(self):
        _check(lib().hd_context_synchronize(self._h))

    def write_flag(self, flag_ptr: int, value: int):
        """enqueue *flag = value on the context's stream (stream memory operation, hd_stream_write_flag)"""
        _check(lib().hd_stream_write_flag(self._h, c_void_p(flag_ptr), int(value)))

    def wait_flag(self, flag_ptr: int, value: int):
        """make the context's stream wait until *flag >= value (stream memory operation, hd_stream_wait_flag)"""
        _check(lib().hd_stream_wait_flag(self._h, c_void_p(flag_ptr), int(value)))

    def timer_start(self):
        _check(lib().hd_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = c_double()
        _check(lib().hd_timer_stop(self._h, byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            lib().hd_context_destroy(self._h)
            self._h = c_void_p()


class MatrixFree:
    """Cartesian phase-space lattice + basis (hyperdeal::MatrixFree::reinit, matrix_free.templates.h:862)."""

    def __init__(self, ctx: Context, dim_x, dim_v, degree, n_cells, left, right, periodic=True, n_points=None, collocation=False,
                 dtype=np.float64, n_cells_global=None, cell_offset=None, side_kind=None):
        dim = dim_x + dim_v
        self.ctx, self.dim_x, self.dim_v, self.dim, self.degree = ctx, dim_x, dim_v, dim, degree
        self.dtype = np.dtype(dtype)
        d = MeshDesc()
        d.dim_x, d.dim_v, d.degree = dim_x, dim_v, degree
        d.n_points = n_points if n_points is not None else degree + 1
        d.collocation = int(bool(collocation))
        d.number_type = HD_F64 if self.dtype == np.float64 else HD_F32
        if isinstance(periodic, bool):
            periodic = (periodic,) * dim
        for i in range(HD_MAX_DIM):
            d.left[i] = left[i] if i < dim else 0.0
            d.right[i] = right[i] if i < dim else 1.0
            d.n_cells[i] = n_cells[i] if i < dim else 1
            d.n_cells_global[i] = (n_cells_global[i] if n_cells_global is not None else n_cells[i]) if i < dim else 1
            d.cell_offset[i] = (cell_offset[i] if cell_offset is not None else 0) if i < dim else 0
            for s in range(2):
                if side_kind is not None and i < dim:
                    d.side_kind[i][s] = side_kind[i][s]
                else:
                    d.side_kind[i][s] = SIDE_PERIODIC_LOCAL if (i >= dim or periodic[i]) else SIDE_DIRICHLET
        self.desc = d
        self.n_cells = tuple(n_cells[:dim])
        self._h = c_void_p()
        _check(lib().hd_mesh_create(ctx._h, byref(d), byref(self._h)))
        self.n_dofs = lib().hd_mesh_n_dofs(self._h)
        self.n_cells_total = lib().hd_mesh_n_cells(self._h)
        self.dofs_per_cell = lib().hd_mesh_dofs_per_cell(self._h)
        self.halo_total = lib().hd_halo_total(self._h)
        self.n_dofs_x = lib().hd_mesh_n_dofs_x(self._h)

    # -- initialize_dof_vector (matrix_free.templates.h:1369)
    def initialize_dof_vector(self, do_ghosts=False) -> int:
        p = c_void_p()
        _check(lib().hd_vector_alloc(self._h, int(do_ghosts), byref(p)))
        return p.value

    def initialize_dof_vector_x(self) -> int:
        """x-space vector (matrix_free_x.initialize_dof_vector of examples/vlasov_poisson/include/application.h)"""
        p = c_void_p()
        _check(lib().hd_vector_alloc_x(self._h, byref(p)))
        return p.value

    def free_vector(self, ptr: int):
        _check(lib().hd_vector_free(self._h, c_void_p(ptr)))

    def copy_in(self, ptr: int, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        _check(lib().hd_vector_copy_in(self._h, c_void_p(ptr), host.ctypes.data_as(c_void_p), host.size))

    def copy_out(self, ptr: int, n: int | None = None) -> np.ndarray:
        out = np.empty(self.n_dofs if n is None else n, dtype=self.dtype)
        _check(lib().hd_vector_copy_out(self._h, c_void_p(ptr), out.ctypes.data_as(c_void_p), out.size))
        return out

    def basis(self, which: int) -> np.ndarray:
        n = _check(lib().hd_mesh_basis(self._h, which, None))
        out = np.empty(n)
        _check(lib().hd_mesh_basis(self._h, which, out.ctypes.data_as(c_void_p)))
        return out

    def halo_offset(self, d, side):
        return lib().hd_halo_offset(self._h, d, side)

    def ghost_size(self, d, side):
        return lib().hd_mesh_ghost_size(self._h, d, side)

    def halo_pack(self, src_ptr: int, send_ptr: int, send_mask=None, peer_dst=None, started_ptr: int | None = None) -> int:
        """send_mask / peer_dst: sequences of 2*HD_MAX_DIM entries indexed 2*dir+side (hd_halo_pack_ex);
        started_ptr: device int every pack CTA increments at start.  Returns the number of CTAs launched."""
        if send_mask is None and peer_dst is None and started_ptr is None:
            _check(lib().hd_halo_pack(self._h, c_void_p(src_ptr), c_void_p(send_ptr)))
            return 0
        mask = (c_int * (2 * HD_MAX_DIM))(*[int(x) for x in send_mask]) if send_mask is not None else None
        dst = (c_void_p * (2 * HD_MAX_DIM))(*[c_void_p(x or 0) for x in peer_dst]) if peer_dst is not None else None
        n = c_int()
        _check(lib().hd_halo_pack_ex(self._h, c_void_p(src_ptr), c_void_p(send_ptr or 0), mask, dst, c_void_p(started_ptr or 0), byref(n)))
        return n.value

    def close(self):
        if self._h:
            lib().hd_mesh_destroy(self._h)
            self._h = c_void_p()


class AdvectionOperation:
    """advection::AdvectionOperation with ConstantVelocityFieldView (advection_operation.h:56-209)."""

    def __init__(self, matrix_free: MatrixFree, velocity, skew_factor: float = 0.0):
        self.mf = matrix_free
        v = (c_double * HD_MAX_DIM)(*([float(x) for x in velocity] + [0.0] * (HD_MAX_DIM - len(velocity))))
        self._h = c_void_p()
        _check(lib().hd_advection_create(matrix_free._h, float(skew_factor), v, byref(self._h)))

    def set_phase_space_velocity(self, a_v_ptr: int | None):
        """a_x = v(q_v), a_v = device table [x-cell][q_x][dim_v] (hd_advection_set_phase_space_velocity); None: constant velocity"""
        _check(lib().hd_advection_set_phase_space_velocity(self._h, c_void_p(a_v_ptr or 0)))

    def apply(self, dst: int, src: int, time: float = 0.0, ghosts: int | None = None):
        """dst = M^-1 A(src, time); dst/src are device pointers (advection_operation.h:137)."""
        _check(lib().hd_advection_apply(self._h, c_void_p(dst), c_void_p(src), c_void_p(ghosts or 0), float(time)))

    def apply_part(self, dst: int, src: int, time: float, ghosts: int | None, part: int):
        """PART_INTERIOR (no ghost data read) / PART_BOUNDARY / PART_ALL: hd_advection_apply_part."""
        _check(lib().hd_advection_apply_part(self._h, c_void_p(dst), c_void_p(src), c_void_p(ghosts or 0), float(time), int(part)))

    @property
    def n_ctas(self) -> int:
        return lib().hd_advection_n_ctas(self._h)

    @property
    def n_halo_senders(self) -> int:
        """CTAs that send the halo in apply_overlapped = increments per arrival counter per application"""
        return lib().hd_advection_n_halo_senders(self._h)

    def set_halo_senders(self, n: int):
        _check(lib().hd_advection_set_halo_senders(self._h, int(n)))

    def apply_overlapped(self, dst: int, src: int, time: float, ghosts: int, sends, counters_ptr: int, target: int):
        """operator + ghost exchange in one kernel (hd_advection_apply_overlapped); sends = [(dir, side, dst_ptr, counter_ptr), ...]"""
        arr = (HaloSend * max(len(sends), 1))()
        for i, (d, s, dp, cp) in enumerate(sends):
            arr[i].dir, arr[i].side, arr[i].dst, arr[i].arrival_counter = int(d), int(s), c_void_p(dp), c_void_p(cp)
        _check(lib().hd_advection_apply_overlapped(self._h, c_void_p(dst), c_void_p(src), c_void_p(ghosts), float(time), arr, len(sends), c_void_p(counters_ptr), int(target)))

    def overlap_timed_out(self) -> bool:
        v = c_int()
        _check(lib().hd_advection_overlap_status(self._h, byref(v)))
        return bool(v.value)

    def ghost_sides(self):
        """needed[2*dir+side]: which ghost sides the operator reads (upwind sides only)."""
        out = (c_int * (2 * HD_MAX_DIM))()
        _check(lib().hd_advection_ghost_sides(self._h, out))
        return list(out)

    def apply_host(self, dst: np.ndarray, src: np.ndarray, time: float = 0.0):
        assert dst.flags.c_contiguous and src.flags.c_contiguous
        _check(lib().hd_advection_apply_host(self._h, dst.ctypes.data_as(c_void_p), src.ctypes.data_as(c_void_p), float(time)))

    def apply_host_ptr(self, dst_ptr: int, src_ptr: int, time: float = 0.0):
        _check(lib().hd_advection_apply_host(self._h, c_void_p(dst_ptr), c_void_p(src_ptr), float(time)))

    def set_evaluation_level(self, level: int):
        """AdvectionOperationEvaluationLevel (advection_operation.h:37-42): EVAL_ALL, EVAL_CELL, EVAL_ALL_WITHOUT_NEIGHBOR_LOAD"""
        _check(lib().hd_advection_set_evaluation_level(self._h, int(level)))

    def set_kernel(self, which: int):
        _check(lib().hd_advection_set_kernel(self._h, which))

    def set_l2_hints(self, mask: int):
        _check(lib().hd_advection_set_l2_hints(self._h, mask))

    def set_row_tile(self, tile):
        """rows per tile along directions 1..5 of the pipelined kernel's cell traversal (hd_advection_set_row_tile)"""
        _check(lib().hd_advection_set_row_tile(self._h, (c_int * 5)(*[int(x) for x in tile])))

    @property
    def kernel_name(self) -> str:
        return lib().hd_advection_kernel_name(self._h).decode()

    @property
    def launch_count(self) -> int:
        return lib().hd_advection_launch_count(self._h)

    def set_dirichlet_builtin(self, fn_id: int):
        _check(lib().hd_advection_set_dirichlet_builtin(self._h, fn_id))

    def set_dirichlet_values(self, d: int, side: int, g: np.ndarray):
        g = np.ascontiguousarray(g, dtype=np.float64)
        _check(lib().hd_advection_set_dirichlet_values(self._h, d, side, g.ctypes.data_as(c_void_p), g.size))

    def close(self):
        if self._h:
            lib().hd_advection_destroy(self._h)
            self._h = c_void_p()


class LowStorageRungeKuttaIntegrator:
    """base/time_integrators.h:48; perform_time_step uses the fused device path (hd_lsrk_step)."""

    def __init__(self, matrix_free: MatrixFree, vec_Ki: int, vec_Ti: int, rk_type: str = "rk45"):
        self.mf, self.Ki, self.Ti = matrix_free, vec_Ki, vec_Ti
        self._h = c_void_p()
        _check(lib().hd_lsrk_create(matrix_free._h, rk_type.encode(), byref(self._h)))

    def n_stages(self) -> int:
        return lib().hd_lsrk_n_stages(self._h)

    def coefficients(self):
        s = self.n_stages()
        b, a = np.empty(s), np.empty(max(s - 1, 1))
        lib().hd_lsrk_coefficients(self._h, 0, b.ctypes.data_as(c_void_p))
        lib().hd_lsrk_coefficients(self._h, 1, a.ctypes.data_as(c_void_p))
        return b, a[: s - 1]

    def perform_time_step(self, solution: int, current_time: float, time_step: float, op):
        """op: an AdvectionOperation (fused device path) or a callable op(src, dst, time) on device
        pointers (unfused path, the reference's std::function signature, time_integrators.h:66-72)."""
        if isinstance(op, AdvectionOperation):
            _check(lib().hd_lsrk_step(self._h, op._h, c_void_p(solution), c_void_p(self.Ki), c_void_p(self.Ti), float(current_time), float(time_step)))
            return
        b, a = self.coefficients()
        L, mf = lib(), self.mf
        _check(L.hd_vector_copy(mf._h, c_void_p(self.Ti), c_void_p(solution)))  # only_Ti_is_ghosted branch
        sum_prev_b = 0.0
        for stage in range(len(b)):
            c = 0.0
            if stage > 0:
                c = sum_prev_b + a[stage - 1]
                sum_prev_b += b[stage - 1]
            op(self.Ti, self.Ki, current_time + c * time_step)
            fa = 0.0 if stage == len(b) - 1 else a[stage] * time_step
            _check(L.hd_lsrk_stage_update(mf._h, c_void_p(solution), c_void_p(self.Ti), c_void_p(self.Ki), b[stage] * time_step, fa))

    def perform_time_step_staged(self, solution: int, current_time: float, time_step: float, op, prepare):
        """Fused stages with a right-hand side that depends on the stage vector (Vlasov-Poisson): before every stage
        `prepare(ti_ptr, stage_time)` refreshes the operator's velocity field from the current stage vector (rho -> Poisson ->
        grad(phi) table), then ONE kernel applies the operator and the stage update (hd_lsrk_stage_fused).  Same numbers as
        perform_time_step with a callable; one streaming kernel less per stage."""
        L = lib()
        b, a = self.coefficients()
        _check(L.hd_vector_copy(self.mf._h, c_void_p(self.Ti), c_void_p(solution)))
        cur, nxt = self.Ti, self.Ki
        sum_prev_b = 0.0
        for stage in range(len(b)):
            c = 0.0
            if stage > 0:
                c = sum_prev_b + a[stage - 1]
                sum_prev_b += b[stage - 1]
            prepare(cur, current_time + c * time_step)
            _check(L.hd_lsrk_stage_fused(self._h, op._h, stage, c_void_p(solution), c_void_p(cur), c_void_p(nxt), None, float(current_time), float(time_step)))
            cur, nxt = nxt, cur

    def perform_time_step_partitioned(self, solution: int, current_time: float, time_step: float, op, peer, ctx):
        """One time step of a brick of a multi-GPU lattice (one process per GPU): per stage ONE kernel
        (hd_lsrk_stage_overlapped) that packs and sends the boundary layers of the current Ti over NVLink, applies the
        operator and the stage update; `peer` is the partition.PeerHaloExchange of this brick.  The reference does the
        same per stage through update_ghost_values + the ECL loop + the update loops (time_integrators.templates.h:93-184)."""
        L = lib()
        _check(L.hd_vector_copy(self.mf._h, c_void_p(self.Ti), c_void_p(solution)))
        cur, nxt = self.Ti, self.Ki
        for stage in range(self.n_stages()):
            g, sends, counters, target = peer.begin_fused(ctx, op)
            arr = (HaloSend * max(len(sends), 1))()
            for i, (d, s, dp, cp) in enumerate(sends):
                arr[i].dir, arr[i].side, arr[i].dst, arr[i].arrival_counter = int(d), int(s), c_void_p(dp), c_void_p(cp)
            _check(L.hd_lsrk_stage_overlapped(self._h, op._h, stage, c_void_p(solution), c_void_p(cur), c_void_p(nxt), c_void_p(g.data_ptr()), arr, len(sends),
                                              c_void_p(counters), int(target), float(current_time), float(time_step)))
            peer.consumed(ctx)
            cur, nxt = nxt, cur

    def close(self):
        if self._h:
            lib().hd_lsrk_destroy(self._h)
            self._h = c_void_p()


class PoissonSolver:
    """x-space field solve of the Vlasov-Poisson right-hand side (hd_poisson_*; examples/vlasov_poisson/include/poisson.h,
    application.h:529-583): density -> potential -> grad(phi) table for AdvectionOperation.set_phase_space_velocity."""

    def __init__(self, matrix_free: MatrixFree):
        self.mf = matrix_free
        self._h = c_void_p()
        _check(lib().hd_poisson_create(matrix_free._h, byref(self._h)))

    def solve(self, rho_x: int, a_v: int, rel_tol: float = 1e-7, max_iterations: int = 10000) -> int:
        """CG steps taken; raises HdError (HD_ERR_NO_CONVERGENCE) if rel_tol is not reached — like the reference's SolverCG"""
        it = c_int()
        _check(lib().hd_poisson_solve(self._h, c_void_p(rho_x), c_void_p(a_v), float(rel_tol), int(max_iterations), byref(it)))
        return it.value

    @property
    def last_solve(self):
        """(iterations, relative residual) of the last solve"""
        it, res = c_int(), c_double()
        _check(lib().hd_poisson_last_solve(self._h, byref(it), byref(res)))
        return it.value, res.value

    @property
    def potential(self) -> int:
        return lib().hd_poisson_potential(self._h)

    def close(self):
        if self._h:
            lib().hd_poisson_destroy(self._h)
            self._h = c_void_p()


class VectorTools:
    @staticmethod
    def interpolate(matrix_free: MatrixFree, vec: int, fn_id: int = FN_HYPERRECTANGLE, time: float = 0.0):
        _check(lib().hd_interpolate_builtin(matrix_free._h, c_void_p(vec), fn_id, float(time)))

    @staticmethod
    def velocity_space_integration(matrix_free: MatrixFree, dst_x: int, src: int):
        """particle density at the x-space nodes (numerics/vector_tools.h:238-315, quad_no_v = 2)"""
        _check(lib().hd_velocity_space_integration(matrix_free._h, c_void_p(dst_x), c_void_p(src)))

    @staticmethod
    def phase_space_diagnostics(matrix_free: MatrixFree, vec: int):
        """[mass, L2 norm, kinetic energy, momentum...] (examples/vlasov_poisson/include/diagnostics.h:34-86)"""
        out = (c_double * 6)()
        _check(lib().hd_phase_space_diagnostics(matrix_free._h, c_void_p(vec), out))
        r = list(out)
        r[1] = math.sqrt(r[1])
        return r

    @staticmethod
    def field_energy(matrix_free: MatrixFree, a_v: int):
        """sum_q (d_d phi)^2 JxW per x-direction (diagnostics.h:88-143) from the gradient table of PoissonSolver.solve"""
        out = (c_double * 3)()
        _check(lib().hd_field_energy(matrix_free._h, c_void_p(a_v), out))
        return list(out)[: matrix_free.dim_x]

    @staticmethod
    def norm_and_error_sums(matrix_free: MatrixFree, vec: int, fn_id: int = FN_HYPERRECTANGLE, time: float = 0.0):
        out = (c_double * 2)()
        _check(lib().hd_norm_and_error_builtin(matrix_free._h, c_void_p(vec), fn_id, float(time), out))
        return out[0], out[1]

    @staticmethod
    def norm_and_error(matrix_free: MatrixFree, vec: int, fn_id: int = FN_HYPERRECTANGLE, time: float = 0.0):
        n2, e2 = VectorTools.norm_and_error_sums(matrix_free, vec, fn_id, time)
        return math.sqrt(n2), math.sqrt(e2)


class MultiGpu:
    """Single-process multi-GPU lattice (hd_multi_*): the phase-space lattice is cut into a Cartesian grid of bricks, one per
    GPU (the reference's PartitionX x PartitionV process grid, examples/advection/include/application.h:150-176 +
    base/mpi.h create_rectangular_comm); ghost faces move over NVLink peer stores.  Vectors are tuples of per-brick
    device pointers; copy_in/copy_out take the GLOBAL lattice in the single-brick ordering."""

    def __init__(self, n_gpus, dim_x, dim_v, degree, n_cells_global, left, right, grid, periodic=True, dtype=np.float64, devices=None):
        dim = dim_x + dim_v
        self.n, self.dim, self.dtype = n_gpus, dim, np.dtype(dtype)
        d = MeshDesc()
        d.dim_x, d.dim_v, d.degree, d.n_points, d.collocation = dim_x, dim_v, degree, degree + 1, 0
        d.number_type = HD_F64 if self.dtype == np.float64 else HD_F32
        if isinstance(periodic, bool):
            periodic = (periodic,) * dim
        for i in range(HD_MAX_DIM):
            d.left[i] = left[i] if i < dim else 0.0
            d.right[i] = right[i] if i < dim else 1.0
            d.n_cells[i] = d.n_cells_global[i] = n_cells_global[i] if i < dim else 1
            d.cell_offset[i] = 0
            for s in range(2):
                d.side_kind[i][s] = SIDE_PERIODIC_LOCAL if (i >= dim or periodic[i]) else SIDE_DIRICHLET
        g = (c_int * HD_MAX_DIM)(*[grid[i] if i < dim else 1 for i in range(HD_MAX_DIM)])
        dev = (c_int * n_gpus)(*devices) if devices is not None else None
        self._h = c_void_p()
        _check(lib().hd_multi_create(n_gpus, dev, byref(d), g, byref(self._h)))
        self.n_dofs = lib().hd_multi_n_dofs(self._h)
        self._ops, self._rks = [], []

    def _arr(self, ptrs):
        return (c_void_p * self.n)(*ptrs)

    def initialize_dof_vector(self):
        a = (c_void_p * self.n)()
        _check(lib().hd_multi_vector_alloc(self._h, a))
        return tuple(a)

    def free(self, vec):
        _check(lib().hd_multi_vector_free(self._h, self._arr(vec)))

    def copy_in(self, vec, host):
        h = np.ascontiguousarray(host, dtype=self.dtype)
        assert h.size == self.n_dofs
        _check(lib().hd_multi_vector_copy_in(self._h, self._arr(vec), h.ctypes.data_as(c_void_p)))

    def copy_out(self, vec):
        h = np.empty(self.n_dofs, dtype=self.dtype)
        _check(lib().hd_multi_vector_copy_out(self._h, self._arr(vec), h.ctypes.data_as(c_void_p)))
        return h

    def interpolate(self, vec, fn_id=FN_HYPERRECTANGLE, time=0.0):
        _check(lib().hd_multi_interpolate_builtin(self._h, self._arr(vec), fn_id, float(time)))

    def norm_and_error(self, vec, fn_id=FN_HYPERRECTANGLE, time=0.0):
        out = (c_double * 2)()
        _check(lib().hd_multi_norm_and_error_builtin(self._h, self._arr(vec), fn_id, float(time), out))
        return math.sqrt(out[0]), math.sqrt(out[1])

    def advection(self, velocity, skew=0.5):
        h = c_void_p()
        v = (c_double * HD_MAX_DIM)(*[velocity[i] if i < self.dim else 0.0 for i in range(HD_MAX_DIM)])
        _check(lib().hd_multi_advection_create(self._h, float(skew), v, byref(h)))
        self._ops.append(h)
        return h

    def kernel_name(self, op):
        return lib().hd_multi_advection_kernel_name(op).decode()

    def apply(self, op, dst, src, time=0.0):
        _check(lib().hd_multi_advection_apply(op, self._arr(dst), self._arr(src), float(time)))

    def lsrk(self, rk_type="rk45"):
        h = c_void_p()
        _check(lib().hd_multi_lsrk_create(self._h, rk_type.encode(), byref(h)))
        self._rks.append(h)
        return h

    def lsrk_step(self, rk, op, solution, Ki, Ti, t, dt):
        _check(lib().hd_multi_lsrk_step(rk, op, self._arr(solution), self._arr(Ki), self._arr(Ti), float(t), float(dt)))

    def synchronize(self):
        _check(lib().hd_multi_synchronize(self._h))

    def close(self):
        if self._h:
            for h in self._rks:
                lib().hd_multi_lsrk_destroy(h)
            for h in self._ops:
                lib().hd_multi_advection_destroy(h)
            lib().hd_multi_destroy(self._h)
            self._h, self._ops, self._rks = c_void_p(), [], []
