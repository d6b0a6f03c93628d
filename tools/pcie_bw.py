import torch, time
n = 403 * 2**20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
h2d = t(lambda: d_in.copy_(h_in, non_blocking=True))
d2h = t(lambda: h_out.copy_(d_out, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
bi = t(both)
print(f"H2D {n/h2d/1e9:.1f} GB/s, D2H {n/d2h/1e9:.1f} GB/s, simultaneous {n/bi/1e9:.1f} GB/s per direction")
