import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
out = torch.zeros(296 * 2, dtype=torch.int64, device="cuda")
for threads in (128,):
    for mode in (0, 8, 256, 257, 264, 265):
        iters = 480
        rc = lib.stk_debug_mma_rate(148, threads, iters, 256, 0, mode, ctypes.c_void_p(out.data_ptr()))
        o = out.cpu().view(-1, 2)[:148]
        print(f"threads={threads} mode={mode}: rc={rc} issue {o[:,0].float().mean()/iters:.1f} clk/MMA, complete {o[:,1].float().mean()/iters:.1f} clk/MMA")
