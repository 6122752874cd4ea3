"""A/B timing of the 3D3V degree-3 FP64 kernels on the 8^6-cell bench lattice: the two-role pipelined kernel against the
three-round kernel and its knobs (library variants from tools/build_variants.sh, HD_R6_PREFETCH, row tiles).  Every
configuration runs in its own process (own CUDA context, 90 s limit), rounds are interleaved so that clock drift hits all
alike, and every result is compared with the first configuration's (strided sample of dst, relative to max|dst|).
usage: python tools/r6_ab.py [name=ENV1=v,ENV2=v[,lib=variant]] ...      (no arguments: the default set)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from hyperdeal_b200 import api
ctx = api.Context(0)
nc = [int(x) for x in os.environ.get("AB_CELLS", "8,8,8,8,8,8").split(",")]
mf = api.MatrixFree(ctx, 3, 3, 3, nc, (0.0,) * 6, (1.0,) * 6)
n = mf.n_dofs
src = torch.empty(n, dtype=torch.float64, device="cuda"); dst = torch.zeros_like(src)
torch.manual_seed(7)
src.normal_()
vel = tuple(float(x) for x in os.environ.get("AB_VEL", "1.0,0.15,-0.05,0.1,-0.15,0.5").split(","))
op = api.AdvectionOperation(mf, vel, 0.5)
reps = int(os.environ.get("REPS", "10"))
for _ in range(3): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
sample = dst[::1021].cpu()
ref_path = os.environ["AB_REF"]
if os.path.exists(ref_path):
    ref = torch.load(ref_path)
    rel = float((sample - ref).abs().max() / ref.abs().max()) if "AB_VEL" not in os.environ else -1.0
else:
    torch.save(sample, ref_path); rel = 0.0
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("%%.4f %%.1f %%.3e %%s %%d" %% (ms, n / ms / 1e6, rel, op.kernel_name, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
''' % ROOT
DEFAULT = ["pipe=HD_FAST_VARIANT=pipe", "rounds=HD_FAST_VARIANT=rounds", "rounds_nopf=HD_FAST_VARIANT=rounds,HD_R6_PREFETCH=0",
           "rounds_pf45=HD_FAST_VARIANT=rounds,HD_R6_PREFETCH=48", "rounds_unroll=HD_FAST_VARIANT=rounds,lib=r6_unroll"]
specs = sys.argv[1:] or DEFAULT
rounds = int(os.environ.get("ROUNDS", "3"))
ref_path = "/tmp/r6_ab_ref_%d.pt" % os.getpid()
res = {}
for r in range(rounds):
    for spec in specs:
        name, _, rest = spec.partition("=")
        env = dict(os.environ, AB_REF=ref_path)
        for kv in filter(None, rest.split(",")):
            k, _, v = kv.partition("=")
            if k == "lib":
                env["HD_LIBHDGPU"] = os.path.join(ROOT, "hyperdeal_b200", "lib", "variants", "libhdgpu_%s.so" % v)
            else:
                env[k] = v
        out = subprocess.run(["timeout", "-s", "KILL", "90", sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print("round %d %-14s FAILED rc=%d %s" % (r, name, out.returncode, (out.stdout + out.stderr)[-600:].replace("\n", " | ")), flush=True)
            continue
        f = out.stdout.split()
        res.setdefault(name, []).append(float(f[0]))
        print("round %d %-14s %s ms  %s GDoF/s  rel.dev vs first %s  kernel %s  sm_clock_after %s" % (r, name, *f), flush=True)
for name, v in res.items():
    best = min(v); med = sorted(v)[len(v) // 2]
    dofs = 4096.0
    for c in os.environ.get("AB_CELLS", "8,8,8,8,8,8").split(","):
        dofs *= int(c)
    print("%-14s best %.4f ms (%.1f GDoF/s)  median %.4f ms (%.1f GDoF/s)" % (name, best, dofs / best / 1e6, med, dofs / med / 1e6))
if os.path.exists(ref_path):
    os.remove(ref_path)
