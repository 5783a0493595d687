import json
import torch
dev = torch.device("cuda")
def timed(fn, n=100):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)
out = {}
for kb in (256, 512, 640, 768, 1024, 1536, 2048):
    n = kb << 10
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    out[f"h2d_{kb}K"] = timed(lambda: d.copy_(h, non_blocking=True))
    out[f"d2h_{kb}K"] = timed(lambda: h.copy_(d, non_blocking=True))
n = 1 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
for parts in (2, 4, 8):
    c = n // parts
    def split():
        for i in range(parts):
            d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
    out[f"h2d_1M_in_{parts}"] = timed(split)
# raw cudaMemcpyAsync through cuda-python-free path: torch uses cudaMemcpyAsync already; try a non-default stream
s = torch.cuda.Stream()
def on_stream():
    with torch.cuda.stream(s):
        d.copy_(h, non_blocking=True)
out["h2d_1M_side_stream"] = timed(on_stream)
print(json.dumps(out))
