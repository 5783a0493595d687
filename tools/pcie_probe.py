"""Host<->device copy times at the sizes the end-to-end path moves per step (pinned memory, CUDA events)."""
import json
import torch

dev = torch.device("cuda")
out = {}
for nbytes in (1 << 17, 1 << 19, 1 << 20, 1 << 22, 1 << 26):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, n=50):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    def both():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    t_h2d = timed(lambda: d.copy_(h, non_blocking=True))
    t_d2h = timed(lambda: h2.copy_(d2, non_blocking=True))
    t_both = timed(both)
    out[str(nbytes)] = {"h2d_us": round(t_h2d, 2), "d2h_us": round(t_d2h, 2), "both_us": round(t_both, 2),
                        "h2d_gbps": round(nbytes / t_h2d / 1e3, 2), "d2h_gbps": round(nbytes / t_d2h / 1e3, 2)}
print(json.dumps(out))
