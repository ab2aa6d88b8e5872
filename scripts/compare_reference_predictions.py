"""After scripts/gpu_reference_dropin.sh: for every layer type, run the reference's inference.py HERE on the CPU with the
oracle layers (oracle/dryrun_reference.py) on the checkpoint the B200 run wrote, and compare its predictions.npz with the
B200 run's own.  Writes profiles/<tag>_reference_dropin.json."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B2G_REFERENCE", "/root/reference")
run = os.path.join(ROOT, "gpurun_out", "refrun")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {"what": "reference inference.py: B200 drop-in predictions vs CPU oracle-layer predictions, same checkpoint", "layers": {}}
for lt in ("GCN", "GAT", "GIN", "Transformer"):
    ck = os.path.join(run, f"ckpt_{lt}", "best_model.pt")
    mine = os.path.join(run, f"pred_{lt}", "predictions.npz")
    if not (os.path.exists(ck) and os.path.exists(mine)):
        out["layers"][lt] = {"error": "missing checkpoint or predictions from the B200 run"}
        continue
    odir = os.path.join(run, f"pred_{lt}_cpu_oracle")
    r = subprocess.run([sys.executable, "-m", "oracle.dryrun_reference", os.path.join(REF, "inference.py"), "--checkpoint", ck,
                        "--device", "cpu", "--output_dir", odir], cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        out["layers"][lt] = {"error": r.stderr[-300:]}
        continue
    a, b = np.load(mine), np.load(os.path.join(odir, "predictions.npz"))
    res = {}
    for k in a.files:
        if k in b.files and a[k].shape == b[k].shape and a[k].dtype.kind == "f":
            ref = np.abs(b[k]).max()
            res[k] = {"shape": list(a[k].shape), "max_abs_ref": float(ref),
                      "max_rel_err": float(np.abs(a[k].astype(np.float64) - b[k]).max() / max(ref, 1e-30))}
    out["layers"][lt] = res
for f in sorted(os.listdir(run)) if os.path.isdir(run) else []:
    if f == "summaries.txt":
        out["summaries"] = [json.loads(l.split("B2G_DROPIN_SUMMARY ", 1)[1]) for l in open(os.path.join(run, f)) if "B2G_DROPIN_SUMMARY" in l]
dst = os.path.join(ROOT, "profiles", f"{tag}_reference_dropin.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["layers"], indent=1))
print("->", dst)
