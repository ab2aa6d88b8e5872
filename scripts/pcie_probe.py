"""Link floor of the host pipeline (streaming.py): pinned H2D alone, D2H alone, both at once, with the chunk size the pipeline
uses (256 K rows x 512 B = 134 MB) and with one large copy.  Prints GB/s per direction."""
import time
import torch

dev = torch.device("cuda:0")
GB = 5_120_000_000
h_in = torch.empty(GB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(GB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(GB, dtype=torch.uint8, device=dev)
d_out = torch.empty(GB, dtype=torch.uint8, device=dev)
h_in.fill_(1); h_out.fill_(2)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, chunk):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for o in range(0, GB, chunk):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for chunk in (GB, 134_217_728, 33_554_432):
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(a, b, chunk)
        t = min(run(a, b, chunk) for _ in range(3))
        print(f"chunk {chunk / 1e6:8.1f} MB {name:5s}: {t * 1e3:7.2f} ms  {GB / t / 1e9:6.2f} GB/s per direction", flush=True)
