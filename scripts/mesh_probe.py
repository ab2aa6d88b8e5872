"""GPU probe: bench.run_mesh_ingest (device mesh ingest, csrc/mesh.cu) on the cfg4 block without the rest of bench.py."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gnn_bfs_rans_b200 as b2g  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


dims = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (250, 200, 200)
print(json.dumps(bench.run_mesh_ingest(b2g, torch.device("cuda:0"), timed, dims)))
