"""Development: event trace of the role-split actor forward kernel (CTA 0): who waits for whom."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic, _lib
ac = ActorCritic(device="cuda:0", seed=1)
n = 148 * 128 * int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 12
obs = torch.rand((n, 12), device="cuda"); out = torch.empty((n, 2), device="cuda")
tr = torch.zeros((3, 256, 2), dtype=torch.int64, device="cuda")
L = ctypes.CDLL(_lib.LIB_PATH)
L.ss_debug_actor_forward_trace.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
for _ in range(3):
    tr.zero_()
    L.ss_debug_actor_forward_trace(ac.actor.data_ptr(), obs.data_ptr(), out.data_ptr(), n, tr.data_ptr(), None, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = min(t[r, 0, 0] for r in range(3) if t[r, 0, 0] > 0)
names = {0: {9: "K 0 entry/1 init/2 staged/3 drained", 1: "P wait MMA1", 2: "P D1 ready", 3: "P epilogue1 done", 4: "P other tile free", 5: "P handed over"},
         1: {1: "Q wait MMA2", 2: "Q D2 ready", 6: "Q half of D2 done", 7: "Q D2 done", 8: "Q stored"},
         2: {1: "M wait P", 2: "M woke", 3: "M MMA1 issued", 4: "M D2 free", 5: "M MMA2 issued"}}
ev = []
for r in range(3):
    for k in range(256):
        if t[r, k, 0] > 0:
            ev.append((int(t[r, k, 0] - t0), names[r][int(t[r, k, 1]) // 100], int(t[r, k, 1]) % 100))
ev.sort()
for c, what, tile in (ev[:40] + [e for e in ev[40:] if e[1].startswith('K')]):
    print("%7d  %-22s tile %d" % (c, what, tile))
