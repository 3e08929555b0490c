"""H2D / D2H bandwidth from page-locked memory with 1, 2, 4 concurrent streams (is one copy engine the limit?)."""
import time
import torch
n = 128 << 20
for direction in ('h2d', 'd2h'):
    for ns in (1, 2, 4):
        hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
        ds = [torch.empty(n, dtype=torch.uint8, device='cuda') for _ in range(8)]
        streams = [torch.cuda.Stream() for _ in range(ns)]
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(8):
                with torch.cuda.stream(streams[k % ns]):
                    if direction == 'h2d':
                        ds[k].copy_(hs[k], non_blocking=True)
                    else:
                        hs[k].copy_(ds[k], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print('%s %d stream(s): %.1f GB/s' % (direction, ns, 8 * n / dt / 1e9))
# both directions at once
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
ds = [torch.empty(n, dtype=torch.uint8, device='cuda') for _ in range(8)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(4):
        with torch.cuda.stream(s1):
            ds[k].copy_(hs[k], non_blocking=True)
        with torch.cuda.stream(s2):
            hs[4 + k].copy_(ds[4 + k], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print('h2d + d2h together: %.1f GB/s total' % (8 * n / dt / 1e9))
