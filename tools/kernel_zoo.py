"""Every kernel of libhdgpu on a representative lattice: ms per launch, GDoF/s and algorithmic GB/s against the measured
HBM copy peak (MEASURED_PEAKS.json).  Algorithmic bytes per DoF: apply 2 x sizeof, fused LSRK stage 4 x sizeof, LSRK stage
update 4 x sizeof, interpolate 1 x sizeof (write), norm_and_error 1 x sizeof (read), halo pack 2 x sizeof per face value."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperdeal_b200 import api
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
ctx = api.Context(0)
V = (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def report(name, ms, ndofs, bytes_per_dof):
    gb = ndofs * bytes_per_dof / ms / 1e6
    print("%-58s %8.3f ms %8.1f GDoF/s %7.0f GB/s alg. = %4.1f %% of %.0f" % (name, ms, ndofs / ms / 1e6, gb, 100 * gb / PEAK, PEAK), flush=True)

def apply_case(name, dx, dv, k, nc, dtype, kernel=0, skew=0.5, nq=None):
    dim = dx + dv
    mf = api.MatrixFree(ctx, dx, dv, k, nc, (0.0,) * dim, (1.0,) * dim, dtype=dtype, n_points=nq)
    op = api.AdvectionOperation(mf, V[:dim], skew)
    op.set_kernel(kernel)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    src = torch.empty(mf.n_dofs, dtype=tdt, device="cuda"); dst = torch.empty_like(src)
    api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
    ms = timeit(lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0))
    report("%s [%s]" % (name, op.kernel_name), ms, mf.n_dofs, 2 * src.element_size())
    return mf, op, src, dst

sel = os.environ.get("ZOO", "all")
if sel in ("all", "apply"):
    apply_case("apply 3D3V k=3 f64, 8^6 cells", 3, 3, 3, [8] * 6, np.float64)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f64, 8^6 cells, generic kernel", 3, 3, 3, [8] * 6, np.float64, kernel=1)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f64, 8^6 cells, tile kernel", 3, 3, 3, [8] * 6, np.float64, kernel=3)
    torch.cuda.empty_cache()
    apply_case("apply 2D2V k=3 f64, 64x64x32x32 cells (configs[0])", 2, 2, 3, [64, 64, 32, 32], np.float64)
    torch.cuda.empty_cache()
    apply_case("apply 2D2V k=3 f64, 64x64x32x32 cells, generic kernel", 2, 2, 3, [64, 64, 32, 32], np.float64, kernel=1)
    torch.cuda.empty_cache()
    apply_case("apply 2D2V k=3 f32, 64x64x32x32 cells", 2, 2, 3, [64, 64, 32, 32], np.float32)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=5 f32, 6x6x6x4x4x4 cells (configs[2])", 3, 3, 5, [6, 6, 6, 4, 4, 4], np.float32)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f32, 8^6 cells", 3, 3, 3, [8] * 6, np.float32)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f32, 8^6 cells, generic kernel", 3, 3, 3, [8] * 6, np.float32, kernel=1)
    torch.cuda.empty_cache()
    apply_case("apply 1D1V k=3 f64, 8192x8192 cells", 1, 1, 3, [8192, 8192], np.float64)
    torch.cuda.empty_cache()
    apply_case("apply 1D1V k=3 f64, 8192x8192 cells, generic kernel", 1, 1, 3, [8192, 8192], np.float64, kernel=1)
    torch.cuda.empty_cache()
    apply_case("apply 2D2V k=3 n_q=5 f64, 32^4 cells (over-integration)", 2, 2, 3, [32] * 4, np.float64, nq=5)
    torch.cuda.empty_cache()
if sel in ("all", "row"):
    apply_case("apply 3D3V k=3 f64, 8^6 cells, row-persistent tile kernel", 3, 3, 3, [8] * 6, np.float64, kernel=4)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f32, 8^6 cells, row-persistent tile kernel", 3, 3, 3, [8] * 6, np.float32, kernel=4)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=3 f64, 8^6 cells", 3, 3, 3, [8] * 6, np.float64)
    torch.cuda.empty_cache()
if sel in ("apply2d",):
    apply_case("apply 2D2V k=3 f64, 64x64x32x32 cells (configs[0])", 2, 2, 3, [64, 64, 32, 32], np.float64)
if sel in ("dirichlet",):
    # Dirichlet lattices: automatic choice (specialised kernel fed with synthesised ghost traces) against the generic kernel + lifting kernel
    def dirichlet_case(name, dx, dv, k, nc, dtype, kernel):
        dim = dx + dv
        mf = api.MatrixFree(ctx, dx, dv, k, nc, (0.0,) * dim, (1.0,) * dim, periodic=False, dtype=dtype)
        op = api.AdvectionOperation(mf, V[:dim], 0.5)
        op.set_dirichlet_builtin(api.FN_HYPERRECTANGLE)
        op.set_kernel(kernel)
        tdt = torch.float64 if dtype == np.float64 else torch.float32
        src = torch.empty(mf.n_dofs, dtype=tdt, device="cuda"); dst = torch.empty_like(src)
        api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
        ms = timeit(lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.1))
        report("%s [%s]" % (name, op.kernel_name), ms, mf.n_dofs, 2 * src.element_size())
        del src, dst
        op.close(); mf.close()
        torch.cuda.empty_cache()
    dirichlet_case("apply 3D3V k=3 f64, 8^6 cells, Dirichlet on all sides", 3, 3, 3, [8] * 6, np.float64, 0)
    dirichlet_case("apply 3D3V k=3 f64, 8^6 cells, Dirichlet on all sides, generic kernel", 3, 3, 3, [8] * 6, np.float64, 1)
    dirichlet_case("apply 2D2V k=3 f64, 64x64x32x32 cells, Dirichlet on all sides", 2, 2, 3, [64, 64, 32, 32], np.float64, 0)
    dirichlet_case("apply 2D2V k=3 f64, 64x64x32x32 cells, Dirichlet on all sides, generic kernel", 2, 2, 3, [64, 64, 32, 32], np.float64, 1)
if sel in ("levels",):
    # attribution of the operator's time with the reference's evaluation levels (advection_operation.h:37-42)
    for kern, kname in ((0, "three-round kernel"), (2, "two-role kernel")):
        mf = api.MatrixFree(ctx, 3, 3, 3, [8] * 6, (0.0,) * 6, (1.0,) * 6)
        op = api.AdvectionOperation(mf, V, 0.5)
        op.set_kernel(kern)
        src = torch.empty(mf.n_dofs, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
        api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
        for level, lname in ((api.EVAL_ALL, "all"), (api.EVAL_ALL_WITHOUT_NEIGHBOR_LOAD, "all_without_neighbor_load"), (api.EVAL_CELL, "cell")):
            op.set_evaluation_level(level)
            ms = timeit(lambda: op.apply(dst.data_ptr(), src.data_ptr(), 0.0))
            report("apply 3D3V k=3 f64, 8^6 cells, %s, level %s" % (kname, lname), ms, mf.n_dofs, 16)
        del src, dst
        op.close(); mf.close()
        torch.cuda.empty_cache()
if sel in ("tg",):  # the global-memory tile kernel on BASELINE.json configs[2] against the generic kernel
    apply_case("apply 3D3V k=5 f32, 6x6x6x4x4x4 cells (configs[2]), global-memory tile kernel", 3, 3, 5, [6, 6, 6, 4, 4, 4], np.float32, kernel=5)
    torch.cuda.empty_cache()
    apply_case("apply 3D3V k=5 f32, 6x6x6x4x4x4 cells (configs[2]), generic kernel", 3, 3, 5, [6, 6, 6, 4, 4, 4], np.float32, kernel=1)
    torch.cuda.empty_cache()
if sel in ("all", "lsrk"):
    mf, op, src, dst = apply_case("apply 3D3V k=3 f64, 8^6 cells (again)", 3, 3, 3, [8] * 6, np.float64)
    Ki = torch.empty_like(src)
    integ = api.LowStorageRungeKuttaIntegrator(mf, Ki.data_ptr(), dst.data_ptr(), "rk45")
    ms = timeit(lambda: integ.perform_time_step(src.data_ptr(), 0.0, 1e-6, op), reps=4, warm=1)
    report("fused rk45 step = 5 stages (per stage:)", ms / 5, mf.n_dofs, 32)
    unf = lambda s, d, t: op.apply(d, s, t)
    ms = timeit(lambda: integ.perform_time_step(src.data_ptr(), 0.0, 1e-6, unf), reps=4, warm=1)
    report("unfused rk45 step = 5 x (apply + stage update) (per stage:)", ms / 5, mf.n_dofs, 48)
    L = api.lib()
    from ctypes import c_void_p
    ms = timeit(lambda: L.hd_lsrk_stage_update(mf._h, c_void_p(src.data_ptr()), c_void_p(dst.data_ptr()), c_void_p(Ki.data_ptr()), 1e-7, 1e-7))
    report("hd_lsrk_stage_update", ms, mf.n_dofs, 32)
    ms = timeit(lambda: api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0))
    report("hd_interpolate_builtin", ms, mf.n_dofs, 8)
    ms = timeit(lambda: api.VectorTools.norm_and_error_sums(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0))
    report("hd_norm_and_error_builtin", ms, mf.n_dofs, 8)
