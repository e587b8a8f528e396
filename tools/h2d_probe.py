"""Development: host-to-device copy bandwidth from pinned memory, one stream vs two concurrent streams, several sizes."""
import time, torch
dev = torch.device("cuda:0")
for mb in (8, 32, 128, 512):
    n = mb * 1024 * 1024
    h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    d = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    s = [torch.cuda.Stream(dev) for _ in range(2)]
    for k in range(2):
        h[k].fill_(1)
    def run(nstreams, reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(reps):
            for k in range(nstreams):
                with torch.cuda.stream(s[k]):
                    d[k].copy_(h[k], non_blocking=True)
        torch.cuda.synchronize()
        return reps * nstreams * n / (time.perf_counter() - t0) / 1e9
    run(1, 3)
    print("%4d MB  one stream %.1f GB/s   two streams %.1f GB/s" % (mb, run(1), run(2)), flush=True)
    # device-to-host beside it
    def duplex(reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(reps):
            with torch.cuda.stream(s[0]):
                d[0].copy_(h[0], non_blocking=True)
            with torch.cuda.stream(s[1]):
                h[1].copy_(d[1], non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9
    print("         h2d with d2h beside it: %.1f GB/s each way" % duplex(), flush=True)
