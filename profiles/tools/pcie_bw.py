import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    for i in range(4): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print("H2D alone GB/s", 4 * n / dt / 1e9)
    torch.cuda.synchronize(); t = time.time()
    for i in range(4): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print("D2H alone GB/s", 4 * n / dt / 1e9)
    torch.cuda.synchronize(); t = time.time()
    with torch.cuda.stream(s1):
        for i in range(4): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        for i in range(4): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print("H2D with concurrent D2H (1/4 size): H2D GB/s", 4 * n / dt / 1e9)
