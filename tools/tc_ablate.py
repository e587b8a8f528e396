"""Development: which role paces the forward kernel?  Time the kernel with the output warps' layer 3 and / or the producers'
layer-1 conversion switched off (results are then wrong on purpose)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic, _lib
ac = ActorCritic(device="cuda:0", seed=1)
n = 524288
obs = torch.rand((n, 12), device="cuda"); out = torch.empty((n, 2), device="cuda")
L = ctypes.CDLL(_lib.LIB_PATH)
L.ss_debug_actor_forward_trace.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
for dbg, what in ((0, "full kernel"), (1, "output warps skip layer 3"), (2, "producers skip the layer-1 conversion"), (3, "both skipped: MMA issue alone")):
    for _ in range(3):
        L.ss_debug_actor_forward_trace(ac.actor.data_ptr(), obs.data_ptr(), out.data_ptr(), n, None, None, dbg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.ss_debug_actor_forward_trace(ac.actor.data_ptr(), obs.data_ptr(), out.data_ptr(), n, None, None, dbg)
    e1.record(); torch.cuda.synchronize()
    print("%-44s %.1f us per 524,288 rows" % (what, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
