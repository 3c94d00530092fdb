import torch, time
n = 2 * 1024**3 // 4
h = torch.empty(n, dtype=torch.int32, pin_memory=True); h.fill_(1)
d = torch.empty(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(2): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("1 stream 2 GiB copies: %.1f GB/s" % (5 * n * 4 / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("2 streams: %.1f GB/s" % (5 * n * 4 / dt / 1e9))
for chunk_mb in (64, 256):
    c = chunk_mb * 1024 * 1024 // 4
    t0 = time.perf_counter()
    for _ in range(3):
        for o in range(0, n, c): d[o:o + c].copy_(h[o:o + c], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("chunks of %d MiB: %.1f GB/s" % (chunk_mb, 3 * n * 4 / dt / 1e9))
