"""Host <-> device copy rates of this box with 1..N GPUs copying at once (pinned host memory, one process, one stream per
GPU): is the end-to-end ceiling of a multi-GPU run the GPUs' own PCIe links or the host side of the box?
usage: python scripts/gpu_pcie_probe_multi.py [--gpus N] [--mb 64]"""
import argparse, json, time
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--gpus", type=int, default=torch.cuda.device_count()); ap.add_argument("--mb", type=int, default=64)
a = ap.parse_args()
n = a.mb << 20
G = min(a.gpus, torch.cuda.device_count())
dev = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
host_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
host_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(G)]
dev_in = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
st = [[torch.cuda.Stream(device=f"cuda:{g}") for _ in range(2)] for g in range(G)]
def run(gpus, d2h=True, h2d=False, reps=10):
    def once():
        for g in gpus:
            if d2h:
                with torch.cuda.stream(st[g][0]): host_out[g].copy_(dev[g], non_blocking=True)
            if h2d:
                with torch.cuda.stream(st[g][1]): dev_in[g].copy_(host_in[g], non_blocking=True)
    def sync():
        for g in gpus:
            st[g][0].synchronize(); st[g][1].synchronize()
    for _ in range(2): once()
    sync(); t0 = time.perf_counter()
    for _ in range(reps): once()
    sync(); dt = (time.perf_counter() - t0) / reps
    return dt
res = {"mb_per_copy": a.mb, "gpus": G, "d2h_alone_gbs": [], "h2d_alone_gbs": []}
for g in range(G):
    res["d2h_alone_gbs"].append(round(n / run([g]) / 1e9, 1))
    res["h2d_alone_gbs"].append(round(n / run([g], d2h=False, h2d=True) / 1e9, 1))
for k in sorted({1, 2, 4, 8, G}):
    if k > G: continue
    gs = list(range(k))
    dt = run(gs); res[f"d2h_concurrent_{k}_aggregate_gbs"] = round(k * n / dt / 1e9, 1)
    dt = run(gs, d2h=False, h2d=True); res[f"h2d_concurrent_{k}_aggregate_gbs"] = round(k * n / dt / 1e9, 1)
    dt = run(gs, d2h=True, h2d=True); res[f"both_concurrent_{k}_aggregate_gbs_each_direction"] = round(k * n / dt / 1e9, 1)
print(json.dumps(res))
