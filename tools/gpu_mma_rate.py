import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib
L = _lib.lib()
for grid in (1, 32, 148):
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    for reps in (10, 100):
        L.r2l_debug_mma_rate(reps, grid, ctypes.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.r2l_debug_mma_rate(reps, grid, ctypes.c_void_p(out.data_ptr()), None); e1.record(); torch.cuda.synchronize()
        cyc = out.double().mean().item()
        print(f"grid {grid:3d} reps {reps:4d}: {cyc / reps:9.1f} cycles per 48-MMA layer ({cyc / reps / 48:6.1f} per MMA), wall {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
