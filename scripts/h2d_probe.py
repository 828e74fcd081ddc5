import time, torch
dev = torch.device("cuda")
def wall(fn, iters=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / iters * 1e6
for kb in (256, 512, 1024, 1536, 2048, 3072, 4096, 8192):
    n = kb * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    t1 = wall(lambda: (d.copy_(h, non_blocking=True), torch.cuda.current_stream().synchronize()))
    t2 = wall(lambda: (h.copy_(d, non_blocking=True), torch.cuda.current_stream().synchronize()))
    half = n // 2
    t3 = wall(lambda: (d[:half].copy_(h[:half], non_blocking=True), d[half:].copy_(h[half:], non_blocking=True), torch.cuda.current_stream().synchronize()))
    print("%5d KB: H2D %.1f us (%.1f GB/s)  D2H %.1f us  H2D in two halves %.1f us" % (kb, t1, n / t1 / 1e3, t2, t3))
h = torch.empty((1 << 20, 2), dtype=torch.uint8).pin_memory(); d = torch.empty((1 << 20, 2), dtype=torch.uint8, device=dev)
print("2D [2^20,2]: %.1f us" % wall(lambda: (d.copy_(h, non_blocking=True), torch.cuda.current_stream().synchronize())))
