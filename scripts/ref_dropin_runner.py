"""Runs one of the reference's OWN scripts (train.py / inference.py, byte-for-byte unchanged) on the B200 through the
drop-in (gnn_bfs_rans_b200.dropin) and reports what ran: libb2g.so kernel launches, the shared objects loaded, and the
wall time between consecutive optimizer steps (a CUDA-synchronised hook around torch.optim.Adam.step — measurement
only; the reference's files are not touched).

    python scripts/ref_dropin_runner.py <dir with the reference scripts>/train.py --layer_type GAT --epochs 2 ...
The summary is one JSON line prefixed with `B2G_DROPIN_SUMMARY`."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gnn_bfs_rans_b200 import _lib, dropin
    script = os.path.abspath(sys.argv[1])
    os.chdir(os.path.dirname(script))                      # the reference resolves 'OpenFOAM-data' relative to its cwd
    lib = _lib.load()
    steps = []
    orig_step = torch.optim.Adam.step

    def timed_step(self, *a, **k):
        out = orig_step(self, *a, **k)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        steps.append(time.perf_counter())
        return out
    torch.optim.Adam.step = timed_step
    _lib.launch_count_reset()
    t0 = time.perf_counter()
    try:
        dropin.run(script, sys.argv[2:])
    finally:
        torch.optim.Adam.step = orig_step
    wall = time.perf_counter() - t0
    gaps = sorted(b - a for a, b in zip(steps[:-1], steps[1:]))
    # a gap spans everything between two optimizer steps of the reference loop: batch.to(device), forward, loss, backward,
    # clip, Adam (train.py:158-197) — plus, across epochs, the validation pass; the lower quartile excludes those
    summary = {"script": os.path.basename(script), "argv": sys.argv[2:], "wall_s": wall,
               "b2g_launch_count": int(_lib.launch_count()), "optimizer_steps": len(steps),
               "train_step_ms_median": 1e3 * gaps[len(gaps) // 2] if gaps else None,
               "train_step_ms_q25": 1e3 * gaps[len(gaps) // 4] if gaps else None,
               "device": torch.cuda.get_device_name(0) if torch.cuda.is_available() else "cpu",
               "loaded_so": sorted({l.split()[-1] for l in open("/proc/self/maps") if "libb2g" in l}),
               "torch_geometric": sys.modules["torch_geometric"].__version__}
    print("B2G_DROPIN_SUMMARY " + json.dumps(summary))


if __name__ == "__main__":
    main()
