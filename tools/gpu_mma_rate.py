"""Tensor-pipe rate of the chain kernels' MMA shapes (no TMA, no epilogue): form 0 = M128 N256 K16 cta_group::1,
1 = M256 cta_group::2 (two tiles per SM pair), 2 = M128 cta_group::2 (one tile per SM pair); variants see r2l_b200.h."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2l_b200 import _lib
L = _lib.lib()
for form, variants in ((0, (0, 1)), (1, (0, 1)), (2, (0, 3, 5, 7))):
    for variant in variants:
        for grid in (2, 64):
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            reps = 100
            assert L.r2l_debug_mma_rate(form, variant, reps, grid, ctypes.c_void_p(out.data_ptr()), None) == 0, L.r2l_last_error()
            torch.cuda.synchronize()
            cyc = out[:: (2 if form else 1)].double().mean().item()
            print(f"form {form} variant {variant} grid {grid:3d}: {cyc / reps:9.1f} cycles per 48-MMA layer ({cyc / reps / 48:6.1f} per MMA)", flush=True)
