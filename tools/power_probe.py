"""Is the 3D3V degree-3 FP64 operator power-bound?  Runs the apply back to back for a few seconds per configuration while
sampling NVML power draw, SM clock and throttle reasons every 5 ms, for real coefficient data (the bench velocity) and for
velocities of 1e-30 (same instruction stream, same memory traffic, FP64 operands that barely toggle the multipliers).
usage: python tools/power_probe.py [seconds]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml
from hyperdeal_b200 import api
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("power limit %.0f W, max SM clock %d MHz" % (pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1e3, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
ctx = api.Context(0)
mf = api.MatrixFree(ctx, 3, 3, 3, [8] * 6, (0.0,) * 6, (1.0,) * 6)
src = torch.randn(mf.n_dofs, dtype=torch.float64, device="cuda"); dst = torch.zeros_like(src)
BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal", 0x20: "sw_thermal", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
for kname, kernel in (("rounds", 6), ("pipe", 2)):
    for vname, vel in (("bench velocity", (1.0, 0.15, -0.05, 0.1, -0.15, 0.5)), ("1e-30 velocity", (1e-30, 1e-30, -1e-30, 1e-30, -1e-30, 1e-30))):
        op = api.AdvectionOperation(mf, vel, 0.5)
        op.set_kernel(kernel)
        for _ in range(3):
            op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
        torch.cuda.synchronize()
        samples, stop = [], [False]
        def poll():
            while not stop[0]:
                samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                time.sleep(0.005)
        th = threading.Thread(target=poll, daemon=True); th.start()
        n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter()
        while time.perf_counter() - t0 < secs:
            for _ in range(20):
                op.apply(dst.data_ptr(), src.data_ptr(), 0.0)
            n += 20
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        stop[0] = True; th.join()
        ms = e0.elapsed_time(e1) / n
        tail = [s for s in samples if s[0] - samples[0][0] > 0.5 * secs]
        first = [s for s in samples if s[0] - samples[0][0] < 0.15]
        reasons = set()
        for s in samples:
            for b, nme in BITS.items():
                if s[3] & b: reasons.add(nme)
        print("%-6s %-15s %7.3f ms/apply %6.1f GDoF/s | first 150 ms: %4.0f W %4.0f MHz | second half: %4.0f W %4.0f MHz | reasons %s" % (
            kname, vname, ms, mf.n_dofs / ms / 1e6, sum(s[1] for s in first) / max(len(first), 1), sum(s[2] for s in first) / max(len(first), 1),
            sum(s[1] for s in tail) / max(len(tail), 1), sum(s[2] for s in tail) / max(len(tail), 1), sorted(reasons)), flush=True)
        time.sleep(1.0)
