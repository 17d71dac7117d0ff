"""PCIe D2H / H2D rate of this box for the e2e leg's transfer sizes (pinned host memory, CUDA events)."""
import torch
n = 62914560
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(19660800, dtype=torch.uint8, device="cuda"); h2 = torch.empty(19660800, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(reps)]; b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
ms = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 62.9 MB alone: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s")
ms = t(lambda: d2.copy_(h2, non_blocking=True)); print(f"H2D 19.7 MB alone: {ms:.3f} ms = {19660800/ms/1e6:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = t(both); print(f"D2H 62.9 MB with H2D 19.7 MB in the other direction: {ms:.3f} ms per pair = {n/ms/1e6:.1f} GB/s D2H")
