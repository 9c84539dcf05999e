import time, torch, threading
dev = torch.device("cuda:0")
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
x = torch.zeros(1024, device=dev)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def burst(k):
    with torch.cuda.stream(sB):
        t = time.perf_counter()
        for _ in range(k):
            x.add_(1.0)
        t1 = time.perf_counter()
        sB.synchronize()
        return (t1 - t) * 1e6, (time.perf_counter() - t) * 1e6
for _ in range(5): burst(27)
for label, copying in (("idle", False), ("during D2H", True), ("during H2D", "h2d")):
    res = []
    for rep in range(20):
        if copying:
            with torch.cuda.stream(sA):
                for _ in range(4):
                    (h.copy_(d, non_blocking=True) if copying is True else d.copy_(h, non_blocking=True))
            time.sleep(0.002)
        res.append(burst(27))
        torch.cuda.synchronize()
    res.sort(key=lambda r: r[1])
    print(label, "27 launches + sync: enqueue %.0f us, total %.0f us (median), max %.0f" % (res[10][0], res[10][1], res[-1][1]))
    res1 = []
    for rep in range(20):
        if copying:
            with torch.cuda.stream(sA):
                for _ in range(4):
                    (h.copy_(d, non_blocking=True) if copying is True else d.copy_(h, non_blocking=True))
            time.sleep(0.002)
        res1.append(burst(1))
        torch.cuda.synchronize()
    res1.sort(key=lambda r: r[1])
    print(label, " 1 launch  + sync: total %.0f us (median), max %.0f" % (res1[10][1], res1[-1][1]))
