"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* by running the UNMODIFIED reference
(/root/reference/openfoam_loader.py + graph_constructor.py, the latter through the stub
torch_geometric.data.Data in oracle/_stub) in this container.  /root/reference does not exist on
the GPU box, so the outputs are committed as small fixtures:

  tests/golden/shipped_mesh.npz      the loader's arrays exactly as the reference hands them to the
                                     builder (owner, neighbour, cell_centers, internal_mask, n_cells)
  tests/golden/builder_golden.json   SHA-256 / shapes / head+tail of every reference builder output
  tests/golden/toy_golden.json       full outputs for hand-sized meshes (known-answer vectors)
  tests/golden/random_golden.npz     full outputs for seeded random face lists (ragged / duplicate /
                                     out-of-range cases)

Run:  python oracle/make_golden.py        (needs /root/reference; ~15 s)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B2G_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_stub"))
sys.path.insert(0, REF)

import torch  # noqa: E402
from graph_constructor import GraphConstructor  # noqa: E402  (the reference, unmodified)
from openfoam_loader import OpenFOAMLoader  # noqa: E402


def sha(a) -> str:
    if isinstance(a, torch.Tensor):
        a = a.contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe(g) -> dict:
    ei = g.edge_index.contiguous()
    return dict(
        num_nodes=int(g.num_nodes), E=int(ei.shape[1]),
        edge_index_sha256=sha(ei), edge_attr_sha256=sha(g.edge_attr), x_sha256=sha(g.x),
        x_shape=list(g.x.shape),
        head=ei[:, :8].tolist(), tail=ei[:, -8:].tolist(),
        n_self_loops=int((ei[0] == ei[1]).sum()),
    )


def full(g) -> dict:
    return dict(num_nodes=int(g.num_nodes), edge_index=g.edge_index.tolist(),
                edge_attr=g.edge_attr.tolist(), x=g.x.tolist())


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)

    # ---- shipped case through the reference loader, verbatim --------------------------------
    mesh = OpenFOAMLoader(os.path.join(REF, "OpenFOAM-data")).load_mesh()
    owner, neighbour = mesh['owner'], mesh['neighbour']
    np.savez_compressed(
        os.path.join(out, "shipped_mesh.npz"),
        owner=owner, neighbour=neighbour, cell_centers=mesh['cell_centers'],
        internal_mask=mesh['internal_mask'], n_cells=np.int64(mesh['n_cells']))
    # the inputs of the loader's get_cell_centers (openfoam_loader.py:191-227) for the same case: points + ragged faces
    sys.path.insert(0, ROOT)
    from oracle import mesh_oracle
    face_pts, face_off = mesh_oracle.flatten_faces(mesh['faces'])
    np.savez_compressed(os.path.join(out, "shipped_polymesh.npz"), points=mesh['points'], face_pts=face_pts,
                        face_off=face_off)
    gc = GraphConstructor(mesh)
    gold = dict(
        loader=dict(owner_sha256=sha(owner), neighbour_sha256=sha(neighbour),
                    cell_centers_sha256=sha(mesh['cell_centers']),
                    n_cells=int(mesh['n_cells']), n_owner=len(owner), n_neighbour=len(neighbour)),
        build_edge_index=dict(sha256=sha(gc.build_edge_index()),
                              shape=list(gc.build_edge_index().shape)),
        mode_A=describe(gc.build_graph(node_features=mesh['cell_centers'], filter_internal=True,
                                       n_internal_cells=12225)),
        mode_B=describe(gc.build_graph(filter_internal=True)),
        mode_C=describe(gc.build_graph(node_features=mesh['cell_centers'])),
    )
    ei_a = gc.build_graph(node_features=mesh['cell_centers'], filter_internal=True,
                          n_internal_cells=12225).edge_index
    gold['compute_edge_attributes_mode_A_sha256'] = sha(gc.compute_edge_attributes(ei_a))
    with open(os.path.join(out, "builder_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)

    # ---- toy known-answer meshes ------------------------------------------------
    cc = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (5, 5, 5), (9, 9, 9)], dtype=np.float64)
    toy_mesh = dict(owner=np.array([0, 0, 1, 2, 0, 0, 1, 1, 2, 2, 3, 3, 5], dtype=np.int32),
                    neighbour=np.array([1, 2, 3, 3], dtype=np.int32), cell_centers=cc, n_cells=6)
    tg = GraphConstructor(dict(toy_mesh))
    toy = dict(
        mesh={k: (v.tolist() if hasattr(v, 'tolist') else v) for k, v in toy_mesh.items()},
        build_edge_index=tg.build_edge_index().tolist(),
        mode_C=full(tg.build_graph()),
        mode_A_n5=full(tg.build_graph(filter_internal=True, n_internal_cells=5)),
        mode_A_n3=full(tg.build_graph(filter_internal=True, n_internal_cells=3)),
        mode_A_n1=full(tg.build_graph(filter_internal=True, n_internal_cells=1)),
        mode_nofilter_fallback=full(tg.build_graph(filter_internal=True)),
    )
    mb = dict(toy_mesh)
    mb['internal_mask'] = np.array([1, 1, 0, 1, 0, 1], dtype=bool)
    toy['mode_B'] = full(GraphConstructor(mb).build_graph(filter_internal=True))
    toy['mode_B_mask'] = mb['internal_mask'].astype(int).tolist()
    # mode C with n_cells smaller than the largest id -> the range filter (:168-173) fires
    ms = dict(toy_mesh)
    ms['n_cells'] = 4
    ms['cell_centers'] = cc[:4]
    toy['mode_C_ncells4'] = full(GraphConstructor(ms).build_graph())
    # no internal faces at all -> only boundary loops; and nothing at all -> all-node self loops
    me = dict(owner=np.array([1, 1, 3], dtype=np.int32), neighbour=np.array([], dtype=np.int32),
              cell_centers=cc[:5], n_cells=5)
    toy['mode_C_no_internal'] = full(GraphConstructor(me).build_graph())
    toy['mode_A_no_internal_n3'] = full(GraphConstructor(me).build_graph(filter_internal=True, n_internal_cells=3))
    toy['mesh_no_internal'] = dict(owner=[1, 1, 3], neighbour=[], n_cells=5)
    # field_data stacking (:242-256)
    fd = dict(U=np.arange(18, dtype=np.float64).reshape(6, 3), p=np.arange(6, dtype=np.float64),
              nut=np.arange(6, dtype=np.float64) * 2)
    toy['mode_C_fields_x'] = tg.build_graph(field_data=fd).x.tolist()
    with open(os.path.join(out, "toy_golden.json"), "w") as f:
        json.dump(toy, f)

    # ---- seeded random face lists: duplicates, pre-existing self loops, ragged sizes ---------
    rng = np.random.default_rng(20261018)
    rnd = {}
    for case, (n_cells, n_int, n_bnd) in enumerate([(50, 120, 30), (200, 150, 0), (64, 0, 10),
                                                    (1000, 3000, 500), (7, 40, 3)]):
        ow = rng.integers(0, n_cells, size=n_int + n_bnd).astype(np.int32)
        ne = rng.integers(0, n_cells, size=n_int).astype(np.int32)
        ccr = rng.standard_normal((n_cells, 3))
        ccr[::7] = ccr[0]                       # coincident centres -> distance 0 branch (:215)
        mask = rng.random(n_cells) < 0.6
        m = dict(owner=ow, neighbour=ne, cell_centers=ccr, n_cells=n_cells, internal_mask=mask)
        g = GraphConstructor(m)
        n_a = max(1, n_cells // 2)
        for tag, gr in (("C", g.build_graph()),
                        ("A", g.build_graph(filter_internal=True, n_internal_cells=n_a)),
                        ("B", g.build_graph(filter_internal=True))):
            rnd[f"c{case}_{tag}_ei"] = gr.edge_index.numpy()
            rnd[f"c{case}_{tag}_ea"] = gr.edge_attr.numpy()
            rnd[f"c{case}_{tag}_x"] = gr.x.numpy()
            rnd[f"c{case}_{tag}_n"] = np.int64(gr.num_nodes)
        rnd[f"c{case}_owner"], rnd[f"c{case}_neighbour"] = ow, ne
        rnd[f"c{case}_cc"], rnd[f"c{case}_mask"] = ccr, mask
        rnd[f"c{case}_ncells"], rnd[f"c{case}_nA"] = np.int64(n_cells), np.int64(n_a)
        rnd[f"c{case}_bei"] = g.build_edge_index().numpy()
    np.savez_compressed(os.path.join(out, "random_golden.npz"), **rnd)
    print(json.dumps(gold, indent=1)[:1500])


if __name__ == "__main__":
    main()
