import torch
x = torch.empty(537*1024*1024//4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for name, fn in (("fill", lambda: x.fill_(1.0)), ("copy", lambda: y.copy_(x)), ("read(sum)", lambda: x.sum())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = x.numel() * 4 / 1e9 * (2 if name == "copy" else 1)
    print(f"{name}: {ms:.4f} ms -> {gb/ms:.2f} TB/s")
