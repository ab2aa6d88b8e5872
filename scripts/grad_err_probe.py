"""Diagnostic (B200): the actual bf16 gradient errors of every layer kind vs the fp64 oracle — per tensor: relative L2,
max-norm, share of rows off by more than 2e-2 of the tensor's max — and the whole-model errors on the shipped graph
(hidden 128, L=4; cfg1 / cfg2), so that the test gates can be set from evidence."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_layers as T  # noqa: E402


def tensor_err(mine, ref):
    ref = ref.double().cpu()
    d = (mine.double().cpu() - ref).abs()
    mx = float(ref.abs().max())
    rows = d.reshape(d.shape[0], -1).max(1).values if d.dim() > 1 else d
    return {"l2": float(d.norm() / ref.norm().clamp_min(1e-30)), "max": float(d.max() / max(mx, 1e-30)),
            "rows_over_2e-2": float((rows > 2e-2 * mx).double().mean()), "ref_max": mx}


res = {}
for kind in T.KINDS:
    for (N, E, F, C) in [(300, 2500, 64, 64), (1000, 6000, 128, 128), (64, 300, 256, 256), (200, 150, 256, 256), (3000, 20000, 256, 256)]:
        ei = T.multigraph(N, E, N + E)
        m = T.make_layer(kind, F, C, torch.bfloat16)
        torch.manual_seed(7)
        x = torch.randn(N, F).to(torch.bfloat16)
        xg = x.cuda().requires_grad_(True)
        out = m(xg, ei.cuda())
        gout = torch.randn(out.shape).to(torch.bfloat16)
        out.backward(gout.cuda())
        x64 = x.double().requires_grad_(True)
        ref, p = T.oracle_forward(kind, m, x64, ei)
        ref.backward(gout.double())
        xb = x.clone().requires_grad_(True)
        refb, pb = T.oracle_forward(kind, m, xb, ei, torch.bfloat16)
        refb.backward(gout)
        r = {"fwd": tensor_err(out.detach(), ref.detach()), "x": tensor_err(xg.grad, x64.grad),
             "x_cpu_bf16_oracle": tensor_err(xb.grad, x64.grad)}
        for name, par in m.named_parameters():
            if par.grad is not None and p[name].grad is not None and float(p[name].grad.abs().max()) > 0:
                r[name] = tensor_err(par.grad, p[name].grad)
                if pb[name].grad is not None:
                    r[name + "_cpu_bf16_oracle"] = tensor_err(pb[name].grad, p[name].grad)
        res[f"{kind}_{N}_{E}_{F}"] = r
        worst = max((v["l2"], k) for k, v in r.items() if not k.endswith("oracle") and k != "fwd")
        print(f"{kind:15s} N={N:5d} E={E:6d} F={F:4d} fwd max {r['fwd']['max']:.2e} | x l2 {r['x']['l2']:.2e} max {r['x']['max']:.2e} "
              f"rows>2e-2 {r['x']['rows_over_2e-2']:.3f} (cpu-bf16 oracle l2 {r['x_cpu_bf16_oracle']['l2']:.2e}) | worst l2 {worst[0]:.2e} {worst[1]}",
              flush=True)

# whole model on the shipped graph
import gnn_bfs_rans_b200 as b2g  # noqa: E402
from gnn_bfs_rans_b200.flow_model import FlowGNN  # noqa: E402
from oracle import layers_oracle as lo  # noqa: E402
z = np.load(os.path.join(ROOT, "tests", "golden", "shipped_mesh.npz"))
mesh = dict(owner=z['owner'], neighbour=z['neighbour'], cell_centers=z['cell_centers'], n_cells=int(z['n_cells']))
g = b2g.GraphConstructor(mesh).build_graph(node_features=mesh['cell_centers'], filter_internal=True, n_internal_cells=12225)
for lt in ("GCN", "GAT", "GIN", "Transformer"):
    for dtype in (torch.float32, torch.bfloat16):
        for training in (False, True):
            torch.manual_seed(0)
            model = FlowGNN(3, 128, 7, 4, lt, dropout=0.0).cuda().to(dtype)
            model.train(training)
            x = g.x.to(dtype)
            xg = x.cuda().requires_grad_(True)
            out = model(xg, g.edge_index.cuda(), g.edge_attr.cuda())
            out.float().square().mean().backward()
            pn = {k for k, _ in model.named_parameters()}
            p = {k: v.detach().double().cpu().requires_grad_(k in pn) for k, v in model.state_dict().items()}
            x64 = x.double().requires_grad_(True)
            ref = lo.flow_gnn_forward(x64, g.edge_index, p, lt, training=training)
            ref.square().mean().backward()
            r = {"fwd": tensor_err(out.detach(), ref.detach()), "x": tensor_err(xg.grad, x64.grad)}
            scale = max(float(p[n].grad.abs().max()) for n in pn if p[n].grad is not None)
            worst = (0.0, "")
            for name, par in model.named_parameters():
                if par.grad is not None and p[name].grad is not None:
                    e = float((par.grad.double().cpu() - p[name].grad).abs().max()) / max(float(p[name].grad.abs().max()), 1e-4 * scale)
                    worst = max(worst, (e, name))
            r["worst_param"] = {"err": worst[0], "name": worst[1]}
            res[f"model_{lt}_{str(dtype)[6:]}_{'train' if training else 'eval'}"] = r
            print(f"model {lt:12s} {str(dtype)[6:]:9s} {'train' if training else 'eval ':5s} fwd max {r['fwd']['max']:.2e} l2 {r['fwd']['l2']:.2e} | "
                  f"x l2 {r['x']['l2']:.2e} max {r['x']['max']:.2e} rows>2e-2 {r['x']['rows_over_2e-2']:.3f} | worst param {worst[0]:.2e} {worst[1]}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "grad_err_probe.json"), "w"), indent=1)
