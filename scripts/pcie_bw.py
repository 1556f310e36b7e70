"""Host<->device copy rates of this box (pinned memory): H2D alone, D2H alone, both at once, for whole-field and per-slice chunk sizes.
The e2e leg of bench.py moves 1.6 GB each way per step, in 32 slices of 50 MB per direction; this says what the link allows."""
import json, sys, time
import torch
dev = torch.device("cuda", 0)
N = 1610612736
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device=dev)
d_out = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for chunk in (N, N // 32):
    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, N, chunk):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, N, chunk):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for f in fns:
                f()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[f"{name}_chunk{chunk >> 20}MB"] = {"ms": round(dt * 1e3, 2), "GBs_per_direction": round(N / dt / 1e9, 1)}
print(json.dumps(res))
