"""Development: event trace of the layer-1 kernel of the frame-stacked actor (CTA 0)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skillshot_learning_b200 import FrameStackActor, _lib
n = 148 * 128 * (int(sys.argv[1]) if len(sys.argv) > 1 else 7)
fa = FrameStackActor(n, frames=20, device="cuda:0", seed=1, precision="bf16")
fa.push(torch.rand((n, 12), device="cuda"))
out = torch.empty((n, 2), device="cuda")
fa.forward(out=out)
DBG = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tr = torch.zeros((3, 256, 2), dtype=torch.int64, device="cuda")
L = ctypes.CDLL(_lib.LIB_PATH)
L.ss_debug_frames_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
for _ in range(3):
    tr.zero_()
    rc = L.ss_debug_frames_trace(fa.params.data_ptr(), fa.stack.data_ptr(), 20, fa.head, out.data_ptr(), n, fa._ws.data_ptr(),
                                 fa._ws.numel(), tr.data_ptr(), None)
    assert rc == 0
torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = min(t[r, 0, 0] for r in range(3) if t[r, 0, 0] > 0)
names = {0: {9: "K entry/staged", 1: "E wait D          tile", 2: "E D ready         tile", 3: "E tile written    tile"},
         1: {1: "M wait D free     tile", 2: "M stage full      t*4+s", 3: "M stage issued    t*4+s"},
         2: {1: "C copy issued     t*4+s"}}
ev = []
for r in range(3):
    for k in range(256):
        if t[r, k, 0] > 0:
            ev.append((int(t[r, k, 0] - t0), names[r][int(t[r, k, 1]) // 100], int(t[r, k, 1]) % 100))
ev.sort()
for c, what, idx in ev[:int(sys.argv[3]) if len(sys.argv) > 3 else 90]:
    print("%7d  %-26s %d" % (c, what, idx))
m = [c for c, w, i in ev if w.startswith("M stage issued")]
f = [c for c, w, i in ev if w.startswith("M stage full")]
print("dbg=%d: staged at %d; per stage: full->issued %.0f cycles; per tile %.0f cycles" % (
    DBG, [c for c, w, i in ev if w.startswith("K")][1], sum(a - b for a, b in zip(m, f)) / len(m), (m[-1] - m[3]) / (len(m) / 4 - 1)))
