"""A/B timing of libhdgpu variants (tools/build_variants.sh) on the 8^6-cell bench lattice.  Each variant runs in its own
process (HD_LIBHDGPU selects the library), rounds are interleaved so that clock drift hits all variants alike; the
result of every variant is compared with the first one's (checksum of dst)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from hyperdeal_b200 import api
ctx = api.Context(0)
mf = api.MatrixFree(ctx, 3, 3, 3, [8] * 6, (0.0,) * 6, (1.0,) * 6)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
api.VectorTools.interpolate(mf, src.data_ptr(), api.FN_HYPERRECTANGLE, 0.0)
op = api.AdvectionOperation(mf, (1.0, 0.15, -0.05, 0.1, -0.15, 0.5), 0.5)
reps = int(os.environ.get("REPS", "10"))
for _ in range(3): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("%%.4f %%.1f %%.17g %%.17g %%d" %% (ms, n / ms / 1e6, float(dst[::4099].sum()), float(dst.abs().max()), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
''' % ROOT
names = sys.argv[1:]
rounds = int(os.environ.get("ROUNDS", "3"))
res = {n: [] for n in names}
for r in range(rounds):
    for n in names:
        lib = os.path.join(ROOT, "hyperdeal_b200", "lib", "variants", "libhdgpu_%s.so" % n)
        out = subprocess.run(["timeout", "-s", "KILL", "40", sys.executable, "-c", CHILD], env=dict(os.environ, HD_LIBHDGPU=lib), capture_output=True, text=True, timeout=None)
        if out.returncode != 0:
            print(n, "FAILED", out.stderr[-400:]); continue
        f = out.stdout.split()
        res[n].append((float(f[0]), float(f[1]), float(f[2]), float(f[3]), int(f[4])))
        print("round %d %-10s %s ms  %s GDoF/s  checksum %s  max %s  sm_clock_after %s" % (r, n, *f), flush=True)
for n in names:
    if res[n]:
        best = min(x[0] for x in res[n]); med = sorted(x[0] for x in res[n])[len(res[n]) // 2]
        print("%-10s best %.4f ms (%.1f GDoF/s)  median %.4f ms" % (n, best, 1.0737e3 / best, med))
