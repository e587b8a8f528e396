"""Development: event trace of the tensor-core critic gradient kernel (CTA 0): epilogue warp 0 and the MMA warp."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import ActorCritic, _lib
ac = ActorCritic(device="cuda:0", seed=1, update_precision="bf16")
n = 148 * 128 * (int(sys.argv[1]) if len(sys.argv) > 1 else 6)
s = torch.rand((n, 12), device="cuda"); a = torch.rand((n, 2), device="cuda") * 2 - 1; y = -torch.rand(n, device="cuda")
g = torch.empty(36609, device="cuda")
tr = torch.zeros((2, 256, 2), dtype=torch.int64, device="cuda")
ws = ac._workspace_for(n)
L = ctypes.CDLL(_lib.LIB_PATH)
L.ss_debug_critic_grad_trace.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
for _ in range(3):
    tr.zero_()
    L.ss_debug_critic_grad_trace(ac.critic.data_ptr(), s.data_ptr(), a.data_ptr(), y.data_ptr(), n, g.data_ptr(), ws.data_ptr(), ws.numel(), tr.data_ptr(), None)
torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = min(t[r, 0, 0] for r in range(2) if t[r, 0, 0] > 0)
step = ["L1a", "L1b", "L2", "G2a+BXa|G2b+G3", "BXb", "G1"]
ev = []
for r in range(2):
    for k in range(256):
        if t[r, k, 0] > 0:
            code = int(t[r, k, 1]); kind, e = code // 1000, code % 1000
            if kind == 0:
                ev.append((int(t[r, k, 0] - t0), "K " + ["entry", "set up + weights staged", "tiles done", "gradient slice written"][code - 900]))
                continue
            what = {1: "E handed over ->", 2: "E resumed after", 3: "M woke for", 4: "M issued"}[kind]
            ev.append((int(t[r, k, 0] - t0), "%s %s (tile %d)" % (what, step[e % 8], e // 8)))
ev.sort()
prev = 0
show = ev[:140] if len(sys.argv) <= 2 else [x for x in ev if x[1].startswith("K")]
for c, what in show:
    print("%7d  (+%5d)  %s" % (c, c - prev, what)); prev = c
