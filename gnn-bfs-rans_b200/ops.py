"""Thin torch-tensor wrappers over the C ABI (include/b2g.h).  PyTorch is plumbing here: it owns
device memory and the stream; every computation is a libb2g.so kernel.  No CPU fallback: a
non-CUDA tensor raises RuntimeError."""
from __future__ import annotations

import os

import torch

from . import _lib

F32, BF16 = 0, 1


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"b2g: unsupported feature dtype {t.dtype} (float32 or bfloat16)")


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b2g: CUDA tensor required (this is the B200 path; there is no CPU fallback)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


_DUMMY = {}


def _p(t):
    """Device pointer of a tensor argument.  An EMPTY tensor (e.g. the column array of a graph without edges) has no
    storage; the gather kernels prefetch index 0 of their index arrays unconditionally (clamped, branch-free loads), so
    empty arguments point at a small zero-filled dummy buffer on the same device instead of NULL."""
    if t is None:
        return None
    if t.numel() == 0 and t.is_cuda:
        d = _DUMMY.get(t.device)
        if d is None:
            d = _DUMMY[t.device] = torch.zeros(64, dtype=torch.int64, device=t.device)
        return d.data_ptr()
    return t.data_ptr()


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D, unit stride in the last dim, 16-byte aligned rows; copy only if the view is not."""
    if t.dim() != 2:
        raise RuntimeError("b2g: expected a 2-D feature matrix")
    es = t.element_size()
    if t.stride(1) != 1 or (t.stride(0) * es) % 16 or t.data_ptr() % 16 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
        if (t.shape[1] * es) % 16:
            raise RuntimeError(f"b2g: feature width {t.shape[1]} x {es} B is not a multiple of 16 bytes")
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def empty_rows(n: int, width: int, dtype, device) -> torch.Tensor:
    """[n, width] view of a buffer whose row stride is a whole number of 128-byte lines: rows of a GEMM operand that start
    mid-line make every TMA box straddle two lines (measured: the k = 1032 dgrad GEMM ran at 3.3 TB/s instead of 4.8)."""
    es = torch.empty((), dtype=dtype).element_size()
    per = 128 // es if os.environ.get("B2G_PAD_ROWS", "1") != "0" else 1
    ld = (width + per - 1) // per * per
    return torch.empty((n, ld), dtype=dtype, device=device)[:, :width]


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------ K0
def build_edge_index(owner: torch.Tensor, neighbour: torch.Tensor) -> torch.Tensor:
    """graph_constructor.py:28-56 on device.  owner/neighbour: int32 CUDA."""
    _cuda(owner, neighbour)
    lib = _lib.load()
    n_o, n_n = owner.numel(), neighbour.numel()
    if n_n > n_o:
        raise IndexError("index out of bounds: neighbour is longer than owner")  # reference: owner[i], :40
    E = n_o + n_n
    out = torch.empty((2, E), dtype=torch.int64, device=owner.device)
    if E:
        _lib.check(lib.b2g_build_edge_index(_p(owner), _p(neighbour), n_o, n_n, _p(out), _stream()), "build_edge_index")
    return out


def mask_to_map(mask_u8: torch.Tensor):
    _cuda(mask_u8)
    lib = _lib.load()
    n = mask_u8.numel()
    o2n = torch.empty(n, dtype=torch.int32, device=mask_u8.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=mask_u8.device)
    ws = _ws(lib.b2g_mask_to_map_workspace_bytes(n), mask_u8.device)
    _lib.check(lib.b2g_mask_to_map(_p(mask_u8), n, _p(o2n), _p(cnt), _p(ws), _stream()), "mask_to_map")
    return o2n, int(cnt.item())


def build_graph_edges(owner, neighbour, mode: int, old_to_new, n_cells: int, n_nodes: int) -> torch.Tensor:
    """Edge part of GraphConstructor.build_graph (graph_constructor.py:109-187, 220-227)."""
    _cuda(owner, neighbour, old_to_new)
    lib = _lib.load()
    dev = owner.device
    n_o, n_n = owner.numel(), neighbour.numel()
    if n_n > n_o:
        raise IndexError("index out of bounds: neighbour is longer than owner")
    ws = _ws(lib.b2g_build_graph_workspace_bytes(n_o, n_n, n_nodes), dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    st = _stream()
    _lib.check(lib.b2g_build_graph_count(_p(owner), _p(neighbour), n_o, n_n, mode, _p(old_to_new), n_cells,
                                         n_nodes, _p(ws), _p(counts), st), "build_graph_count")
    e_kept, n_iso, n_bad = (int(v) for v in counts.tolist())      # the one host sync of the builder
    if n_bad:
        raise IndexError(f"index out of bounds: {n_bad} internal faces reference cells outside [0, {n_cells})")
    E = e_kept + n_iso
    out = torch.empty((2, E), dtype=torch.int64, device=dev)
    if E:
        _lib.check(lib.b2g_build_graph_fill(_p(owner), _p(neighbour), n_o, n_n, mode, _p(old_to_new), n_cells,
                                            n_nodes, _p(ws), E, _p(out), st), "build_graph_fill")
    return out


def edge_attr(cell_centers_f64: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """graph_constructor.py:58-90 / :190-219 on device.  fp64 [n,3] centres -> fp32 [E,4]."""
    _cuda(cell_centers_f64, edge_index)
    lib = _lib.load()
    cc = cell_centers_f64.contiguous()
    ei = edge_index.contiguous()
    E = ei.shape[1]
    out = torch.empty((E, 4), dtype=torch.float32, device=ei.device)
    if E:
        _lib.check(lib.b2g_edge_attr(_p(cc), cc.shape[0], _p(ei), E, _p(out), _stream()), "edge_attr")
    return out


# ------------------------------------------------------------------------------------------ mesh ingest (§8f-3)
def mesh_num_cells(owner: torch.Tensor, neighbour: torch.Tensor) -> int:
    """openfoam_loader.py:197 on device: max(max(owner), max(neighbour)) + 1 (one host read of the result)."""
    _cuda(owner, neighbour)
    lib = _lib.load()
    if owner.numel() + neighbour.numel() == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # np.max, :197
    out = torch.empty(1, dtype=torch.int64, device=owner.device)
    _lib.check(lib.b2g_mesh_num_cells(_p(owner), owner.numel(), _p(neighbour), neighbour.numel(), _p(out),
                                      _p(_ws(256, owner.device)), _stream()), "mesh_num_cells")
    return int(out)


def mesh_cell_centers(points: torch.Tensor, owner: torch.Tensor, neighbour: torch.Tensor, face_pts: torch.Tensor,
                      face_off: torch.Tensor, n_cells: int, n_slots: int) -> torch.Tensor:
    """openfoam_loader.py:191-227 on device.  points fp64 [P,3]; owner / neighbour / face_pts int32; face_off int64
    [F+1]; n_slots = face_off[len(owner)] + face_off[len(neighbour)].  -> fp64 [n_cells, 3]."""
    _cuda(points, owner, neighbour, face_pts, face_off)
    lib = _lib.load()
    dev = points.device
    out = torch.empty((n_cells, 3), dtype=torch.float64, device=dev)
    bad = torch.empty(1, dtype=torch.int64, device=dev)
    nbytes = lib.b2g_mesh_workspace_bytes(n_cells, n_slots)
    ws = _ws(nbytes, dev)
    _lib.check(lib.b2g_mesh_cell_centers(_p(points), points.shape[0], _p(owner), owner.numel(), _p(neighbour),
                                         neighbour.numel(), _p(face_off), _p(face_pts), face_off.numel() - 1, n_slots,
                                         n_cells, _p(out), _p(bad), _p(ws), ws.numel(), _stream()), "mesh_cell_centers")
    if int(bad):
        raise IndexError(f"b2g.mesh: {int(bad)} cell / vertex ids out of range")        # :205-219 IndexError upstream
    return out


def mesh_internal_cells(owner: torch.Tensor, neighbour: torch.Tensor, n_cells: int) -> torch.Tensor:
    """openfoam_loader.py:229-248 on device -> bool [n_cells]."""
    _cuda(owner, neighbour)
    lib = _lib.load()
    dev = owner.device
    mask = torch.empty(n_cells, dtype=torch.uint8, device=dev)
    bad = torch.empty(1, dtype=torch.int64, device=dev)
    _lib.check(lib.b2g_mesh_internal_cells(_p(owner), owner.numel(), _p(neighbour), neighbour.numel(), n_cells,
                                           _p(mask), _p(bad), _p(_ws(256, dev)), _stream()), "mesh_internal_cells")
    if int(bad):
        raise IndexError(f"b2g.mesh: {int(bad)} cell ids out of range")
    return mask.bool()


# ------------------------------------------------------------------------------------------ K1
def csr_build(edge_index: torch.Tensor, N: int, self_loops: bool, by_source: bool, want_dinv: bool):
    """-> (rowptr int32[N+1], col int32[nnz], eid int32[nnz], dinv fp32[N] | None)."""
    _cuda(edge_index)
    lib = _lib.load()
    ei = edge_index.contiguous()
    dev = ei.device
    E = ei.shape[1]
    ws = _ws(lib.b2g_csr_workspace_bytes(E, N), dev)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    nnz_d = torch.zeros(2, dtype=torch.int64, device=dev)
    st = _stream()
    _lib.check(lib.b2g_csr_count(_p(ei), E, N, int(self_loops), int(by_source), _p(rowptr), _p(nnz_d), _p(ws), st), "csr_count")
    nnz, bad = (int(v) for v in nnz_d.tolist())                   # one host sync per distinct edge_index
    if bad:
        raise RuntimeError(f"b2g: edge_index has {bad} edges with an endpoint outside [0, {N})")
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    eid = torch.empty(nnz, dtype=torch.int32, device=dev)
    dinv = torch.empty(N, dtype=torch.float32, device=dev) if want_dinv else None
    _lib.check(lib.b2g_csr_fill(_p(ei), E, N, int(self_loops), int(by_source), _p(rowptr), nnz, _p(col), _p(eid),
                                _p(dinv), _p(ws), st), "csr_fill")
    return rowptr, col, eid, dinv


def csr_perm(eid_a, eid_b, id_space: int):
    lib = _lib.load()
    nnz = eid_a.numel()
    scratch = torch.empty(max(id_space, 1), dtype=torch.int32, device=eid_a.device)
    perm = torch.empty(max(nnz, 1), dtype=torch.int32, device=eid_a.device)
    _lib.check(lib.b2g_csr_perm(_p(eid_a), _p(eid_b), nnz, _p(scratch), _p(perm), _stream()), "csr_perm")
    return perm


# ------------------------------------------------------------------------------------------ K2/K3
def seg_sum(x, rowptr, col, n_rows, row_scale=None, col_scale=None, self_coef=0.0, x_self=None, bias=None,
            relu=False, out=None, band=0, tune=None):
    """tune = (impl, chunk_rows, panel_rows): per-call kernel / row-schedule choice (tests, probes); results do not
    depend on it."""
    _cuda(x)
    lib = _lib.load()
    x = _rows(x)
    xs = _rows(x_self) if x_self is not None else None
    F = x.shape[1]
    if out is None:
        out = torch.empty((n_rows, F), dtype=x.dtype, device=x.device)
    if tune is not None:
        impl, chunk, panel = tune
        _lib.check(lib.b2g_seg_sum_tuned(_p(x), _ld(x), _p(xs), _ld(xs) if xs is not None else 0, _p(out), _ld(out),
                                         n_rows, F, _dt(x), _p(rowptr), _p(col), _p(row_scale), _p(col_scale),
                                         float(self_coef), _p(bias), int(relu), int(band), 0, int(impl), int(chunk),
                                         int(panel), _stream()), "seg_sum_tuned")
        return out
    _lib.check(lib.b2g_seg_sum_banded(_p(x), _ld(x), _p(xs), _ld(xs) if xs is not None else 0, _p(out), _ld(out),
                                      n_rows, F, _dt(x), _p(rowptr), _p(col), _p(row_scale), _p(col_scale),
                                      float(self_coef), _p(bias), int(relu), int(band), _stream()), "seg_sum")
    return out


def segw_gemm_supported(n: int, F: int, C: int, dtype) -> bool:
    dt = F32 if dtype == torch.float32 else (BF16 if dtype == torch.bfloat16 else -1)
    return dt >= 0 and bool(_lib.load().b2g_segw_gemm_supported(int(max(n, 1)), int(F), int(C), dt))


def segw_gemm(x, rowptr, col, n_rows, w, bias=None, col_scale=None, row_scale=None, self_coef=0.0, relu=False, band=0, out=None):
    """out [n_rows, C] = act(row_scale * (sum_j col_scale_j x_j + self_coef x_i) W^T + bias) in ONE kernel (csrc/gcn_fused.cu):
    the CSR segment-sum feeding a tcgen05 GEMM from shared memory (GCNConv forward; GINConv + first Linear of its MLP)."""
    _cuda(x, w)
    x, w = _rows(x), _rows(w)
    if w.dtype != x.dtype:
        w = w.to(x.dtype)
    C, F = w.shape
    if out is None:
        out = torch.empty((n_rows, C), dtype=x.dtype, device=x.device)
    b = bias.float().contiguous() if bias is not None else None
    _lib.check(_lib.load().b2g_segw_gemm(_p(x), _ld(x), _p(rowptr), _p(col), _p(col_scale), _p(row_scale), float(self_coef),
                                         _p(w), _ld(w), _p(b), int(relu), _p(out), _ld(out), n_rows, F, C, _dt(x), int(band),
                                         _stream()), "segw_gemm")
    return out


def colsum(x) -> torch.Tensor:
    _cuda(x)
    lib = _lib.load()
    x = _rows(x)
    F = x.shape[1]
    out = torch.empty(F, dtype=torch.float32, device=x.device)
    ws = _ws(lib.b2g_colsum_workspace_bytes(F), x.device)
    _lib.check(lib.b2g_colsum(_p(x), _ld(x), x.shape[0], F, _dt(x), _p(out), _p(ws), _stream()), "colsum")
    return out


# ------------------------------------------------------------------------------------------ K6
GEMM_IMPL = 0  # 0 auto, 1 SIMT, 2 tcgen05 (tests force one or the other)


def linear_fwd(x, w, bias=None, row_scale=None, act=0, m_main=None, out=None, reserve_sms=0):
    """y = act(row_scale * (x @ w.T) + bias).  Returns (y [n,m_main], aux fp32 [n,m-m_main] | None).
    reserve_sms: SMs the persistent tensor-core kernel leaves to a collective running on another stream."""
    _cuda(x, w)
    lib = _lib.load()
    x = _rows(x)
    w = _rows(w)
    if w.dtype != x.dtype:
        w = w.to(x.dtype)
    n, k = x.shape
    m = w.shape[0]
    if w.shape[1] != k:
        raise RuntimeError(f"b2g linear: weight [{m},{w.shape[1]}] does not match input width {k}")
    m_main = m if m_main is None else m_main
    y = out if out is not None else torch.empty((n, m_main), dtype=x.dtype, device=x.device)
    aux = torch.empty((n, m - m_main), dtype=torch.float32, device=x.device) if m_main < m else None
    b = bias.float().contiguous() if bias is not None else None
    ws = _ws(lib.b2g_linear_workspace_bytes(n, m, k, _dt(x), 0), x.device)
    _lib.check(lib.b2g_linear_fwd(_p(x), _ld(x), _p(w), _ld(w), _p(b), _p(row_scale), _p(y), _ld(y), _p(aux),
                                  (m - m_main) if aux is not None else 0, n, m, m_main, k, _dt(x), int(act),
                                  GEMM_IMPL | (max(0, min(int(reserve_sms), 128)) << 8), _p(ws), _stream()), "linear_fwd")
    return y, aux


def linear_dgrad(dy, w, out=None):
    """dx[n,k] = dy[n,m] @ w[m,k]."""
    _cuda(dy, w)
    lib = _lib.load()
    dy = _rows(dy)
    w = _rows(w)
    if w.dtype != dy.dtype:
        w = w.to(dy.dtype)
    n, m = dy.shape
    k = w.shape[1]
    dx = out if out is not None else torch.empty((n, k), dtype=dy.dtype, device=dy.device)
    if GEMM_IMPL != 1 and lib.b2g_linear_impl(n, k, m, _dt(dy), 0) == 2:
        # tensor-core path: dgrad is a forward GEMM against W^T (a [k,m] copy of a few hundred KB)
        y, _ = linear_fwd(dy, w.t().contiguous(), out=dx)
        return y
    ws = _ws(256, dy.device)
    _lib.check(lib.b2g_linear_dgrad(_p(dy), _ld(dy), _p(w), _ld(w), _p(dx), _ld(dx), n, m, k, _dt(dy), 1,
                                    _p(ws), _stream()), "linear_dgrad")
    return dx


def linear_dgrad_masked(dy, w, mask):
    """dx = (dy @ w) where mask > 0, else 0: the ReLU backward fused into the dgrad GEMM's epilogue (bf16 tensor-core path);
    returns None when the shape / dtype has no such kernel (the caller masks separately)."""
    _cuda(dy, w, mask)
    lib = _lib.load()
    if dy.dtype != torch.bfloat16 or GEMM_IMPL == 1:
        return None
    dy, w, mask = _rows(dy), _rows(w), _rows(mask)
    n, m = dy.shape
    k = w.shape[1]
    if lib.b2g_linear_impl(n, k, m, _dt(dy), 0) != 2:
        return None
    wt = (w if w.dtype == dy.dtype else w.to(dy.dtype)).t().contiguous()                 # [k, m]
    dx = torch.empty((n, k), dtype=dy.dtype, device=dy.device)
    ws = _ws(lib.b2g_linear_workspace_bytes(n, k, m, _dt(dy), 0), dy.device)
    rc = lib.b2g_linear_fwd_masked(_p(dy), _ld(dy), _p(wt), _ld(wt), _p(mask), _ld(mask), _p(dx), _ld(dx), n, k, m, _dt(dy),
                                   _p(ws), _stream())
    if rc == _lib.E_UNSUPPORTED:
        return None
    _lib.check(rc, "linear_fwd_masked")
    return dx


def linear_wgrad(dy, x, want_bias=True):
    """dW[m,k] = dy.T @ x (fp32), db[m] = dy.sum(0) (fp32)."""
    _cuda(dy, x)
    lib = _lib.load()
    dy = _rows(dy)
    x = _rows(x)
    n, m = dy.shape
    k = x.shape[1]
    dw = torch.empty((m, k), dtype=torch.float32, device=dy.device)
    ws = _ws(lib.b2g_linear_workspace_bytes(n, m, k, _dt(dy), 2), dy.device)
    if GEMM_IMPL != 1 and n > 0 and lib.b2g_linear_impl(n, m, k, _dt(dy), 2) == 2:
        # tcgen05 wgrad; the bias gradient is a separate column-sum kernel
        _lib.check(lib.b2g_linear_wgrad(_p(dy), _ld(dy), _p(x), _ld(x), _p(dw), k, None, n, m, k, _dt(dy), 2,
                                        _p(ws), _stream()), "linear_wgrad(tc)")
        return dw, (colsum(dy) if want_bias else None)
    db = torch.empty(m, dtype=torch.float32, device=dy.device) if want_bias else None
    _lib.check(lib.b2g_linear_wgrad(_p(dy), _ld(dy), _p(x), _ld(x), _p(dw), k, _p(db), n, m, k, _dt(dy), 1,
                                    _p(ws), _stream()), "linear_wgrad")
    return dw, db


# ------------------------------------------------------------------------------------------ K4
def gat_fwd(xw, a, H, C, concat, slope, rowptr, col, bias, p_drop, seed, save_stats, max_degree=0):
    """xw [N,H*C]; a fp32 [N,2H] = [a_src | a_dst].  -> out, smax, ssum."""
    lib = _lib.load()
    N = xw.shape[0]
    out = torch.empty((N, H * C if concat else C), dtype=xw.dtype, device=xw.device)
    smax = torch.empty((N, H), dtype=torch.float32, device=xw.device) if save_stats else None
    ssum = torch.empty((N, H), dtype=torch.float32, device=xw.device) if save_stats else None
    _lib.check(lib.b2g_gat_fwd(_p(xw), _ld(xw), _p(a), a.data_ptr() + 4 * H, a.stride(0), _p(out), _ld(out), N, H, C,
                               _dt(xw), int(concat), float(slope), _p(rowptr), _p(col), _p(bias), _p(smax), _p(ssum),
                               float(p_drop), int(seed), int(max_degree), _stream()), "gat_fwd")
    return out, smax, ssum


def gat_bwd(xw, a, gout, H, C, concat, slope, csr, csr_t, perm, smax, ssum, p_drop, seed, d_xw_out):
    """-> d_a fp32 [N,2H]; writes d_xw into d_xw_out (a [N,H*C] view, any row stride)."""
    lib = _lib.load()
    N = xw.shape[0]
    dev = xw.device
    nnz = csr[1].numel()
    alpha_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    ds_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    d_a = torch.empty((N, 2 * H), dtype=torch.float32, device=dev)
    gout = _rows(gout)
    st = _stream()
    _lib.check(lib.b2g_gat_bwd_dst(_p(xw), _ld(xw), _p(a), a.data_ptr() + 4 * H, a.stride(0), _p(gout), _ld(gout), N,
                                   H, C, _dt(xw), int(concat), float(slope), _p(csr[0]), _p(csr[1]), _p(smax),
                                   _p(ssum), float(p_drop), int(seed), _p(alpha_e), _p(ds_e),
                                   d_a.data_ptr() + 4 * H, 2 * H, st), "gat_bwd_dst")
    _lib.check(lib.b2g_gat_bwd_src(_p(gout), _ld(gout), _p(alpha_e), _p(ds_e), _p(d_xw_out), _ld(d_xw_out),
                                   _p(d_a), 2 * H, N, H, C, _dt(xw), int(concat), _p(csr_t[0]), _p(csr_t[1]),
                                   _p(perm), st), "gat_bwd_src")
    return d_a


# ------------------------------------------------------------------------------------------ K4, aggregate-first
def gatz_supported(n: int, H: int, F: int, dtype) -> bool:
    dt = F32 if dtype == torch.float32 else (BF16 if dtype == torch.bfloat16 else -1)
    return dt >= 0 and bool(_lib.load().b2g_gatz_supported(int(max(n, 1)), int(H), int(F), dt))


def rowdot8(x, V) -> torch.Tensor:
    """a[N,8] (fp32) = x[N,F] @ V[8,F]^T."""
    _cuda(x, V)
    x = _rows(x)
    V = V.float().contiguous()
    N, F = x.shape
    out = torch.empty((N, 8), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b2g_rowdot8(_p(x), _ld(x), _p(V), V.stride(0), _p(out), 8, N, F, _dt(x), _stream()), "rowdot8")
    return out


def gatz_fwd(x, a, H, slope, rowptr, col, p_drop, seed, save_stats, band=0, out=None):
    """z[N, H*F] = per-head attention-weighted sums of the F-wide rows of x; a = [a_src | a_dst] fp32 [N, 2H]."""
    x = _rows(x)
    N, F = x.shape
    z = out if out is not None else torch.empty((N, H * F), dtype=x.dtype, device=x.device)
    smax = torch.empty((N, H), dtype=torch.float32, device=x.device) if save_stats else None
    ssum = torch.empty((N, H), dtype=torch.float32, device=x.device) if save_stats else None
    _lib.check(_lib.load().b2g_gatz_fwd(_p(x), _ld(x), _p(a), a.stride(0), _p(z), _ld(z), N, H, F, _dt(x), float(slope),
                                        _p(rowptr), _p(col), _p(smax), _p(ssum), float(p_drop), int(seed), int(band),
                                        _stream()), "gatz_fwd")
    return z, smax, ssum


def gatw_gemm_supported(n: int, H: int, F: int, C: int, dtype) -> bool:
    dt = F32 if dtype == torch.float32 else (BF16 if dtype == torch.bfloat16 else -1)
    return dt >= 0 and bool(_lib.load().b2g_gatw_gemm_supported(int(max(n, 1)), int(H), int(F), int(C), dt))


def gat_alpha(a, rowptr, col, H, slope, p_drop, seed, save_stats, edge_bias=None, rows=None, alpha_out=None):
    """Post-dropout attention weights alpha fp32 [nnz, H] in target-major CSR order (+ softmax max / sum [N, H] | None)
    from a = [a_src | a_dst] fp32 [N, 2H]  (csrc/gat_fused.cu gat_alpha_kernel)."""
    _cuda(a, rowptr, col)
    N = a.shape[0]
    r0, n = (0, rowptr.numel() - 1) if rows is None else (rows[0], rows[1] - rows[0])      # rows = (r0, r1): a row range
    alpha = alpha_out if alpha_out is not None else torch.empty((max(col.numel(), 1), H), dtype=torch.float32, device=a.device)
    smax = torch.empty((N, H), dtype=torch.float32, device=a.device) if save_stats else None
    ssum = torch.empty((N, H), dtype=torch.float32, device=a.device) if save_stats else None
    _lib.check(_lib.load().b2g_gat_alpha(_p(a), a.stride(0), _p(rowptr), _p(col), _p(edge_bias), r0, n, H, float(slope),
                                         float(p_drop), int(seed), _p(alpha), _p(smax), _p(ssum), _stream()), "gat_alpha")
    return alpha, smax, ssum


def gatw_gemm(x, rowptr, col, perm, alpha, wp, bias, n_rows, H, band=0, out=None, srow=None, bvh=None, addend=None):
    """out [n_rows, C] = (per-head alpha-weighted sums of the rows of x over the CSR) @ Wc^T + bias in ONE kernel
    (csrc/gat_fused.cu); wp = Wc with columns permuted by 64-feature chunk (see include/b2g.h).
    TransformerConv terms (optional): + sum_h srow[i, h] bvh[h, :] (fp32 [n_rows, 4] / [4, C]) + addend[i, :] ([n_rows, C])."""
    _cuda(x, wp, alpha)
    x, wp = _rows(x), _rows(wp)
    C = wp.shape[0]
    F = x.shape[1]
    if out is None:
        out = torch.empty((n_rows, C), dtype=x.dtype, device=x.device)
    b = bias.float().contiguous() if bias is not None else None
    if srow is None and bvh is None and addend is None:
        _lib.check(_lib.load().b2g_gatw_gemm(_p(x), _ld(x), _p(rowptr), _p(col), _p(perm), _p(alpha), _p(wp), _ld(wp), _p(b),
                                             _p(out), _ld(out), n_rows, H, F, C, _dt(x), int(band), _stream()), "gatw_gemm")
        return out
    ad = _rows(addend) if addend is not None else None
    if srow is not None:
        assert srow.dtype == torch.float32 and srow.is_contiguous() and srow.shape[1] == 4 and bvh.dtype == torch.float32 \
            and bvh.is_contiguous() and tuple(bvh.shape) == (4, C)
    _lib.check(_lib.load().b2g_gatw_gemm_ex(_p(x), _ld(x), _p(rowptr), _p(col), _p(perm), _p(alpha), _p(wp), _ld(wp), _p(b),
                                            _p(srow), _p(bvh), _p(ad), _ld(ad) if ad is not None else 0, _p(out), _ld(out),
                                            n_rows, H, F, C, _dt(x), int(band), _stream()), "gatw_gemm_ex")
    return out


def gatw_gemm_sm(x, a, rowptr, col, wp, bias, n_rows, H, slope, p_drop, seed, save_stats, max_row_len, band=0, out=None,
                 edge_bias=None, row0=0):
    """GATConv forward with the softmax inside the fused kernel (every row <= 8 entries; csrc/gat_fused.cu): a = [a_src | a_dst]
    fp32 [N, 2H] indexed globally, rows [row0, row0 + n_rows) of it are this call's targets (rowptr is their slice).
    -> (out [n_rows, C], smax, ssum fp32 [n_rows, H] | None)."""
    _cuda(x, wp, a)
    x, wp = _rows(x), _rows(wp)
    C, F = wp.shape[0], x.shape[1]
    if out is None:
        out = torch.empty((n_rows, C), dtype=x.dtype, device=x.device)
    b = bias.float().contiguous() if bias is not None else None
    smax = torch.empty((n_rows, H), dtype=torch.float32, device=x.device) if save_stats else None
    ssum = torch.empty((n_rows, H), dtype=torch.float32, device=x.device) if save_stats else None
    assert a.dtype == torch.float32 and a.stride(1) == 1 and a.shape[1] >= 2 * H
    a_dst = a[row0:, H:]
    _lib.check(_lib.load().b2g_gatw_gemm_sm(_p(x), _ld(x), _p(rowptr), _p(col), _p(a), _p(a_dst), a.stride(0), _p(edge_bias),
                                            float(slope), float(p_drop), int(seed), _p(smax), _p(ssum), _p(wp), _ld(wp), _p(b),
                                            _p(out), _ld(out), n_rows, int(max_row_len), H, F, C, _dt(x), int(band), _stream()),
               "gatw_gemm_sm")
    return out, smax, ssum


def tz_alpha(x, u, H, rowptr, col, p_drop, seed, band=0, edge_bias=None, impl=0):
    """TransformerConv attention weights without the weighted sums (gat_rows.cu: tz_alpha_mma_kernel for bf16 F = 256 — the
    logits as m16n8k16 tile products —, else / impl=1 tz_fwd_kernel in alpha-only mode):
    -> (alpha_pre [nnz, H], alpha_post | None (p_drop == 0), ssum fp32 [N, H] = per-head sums of the post-dropout weights)."""
    x, u = _rows(x), _rows(u)
    N, F = u.shape[0], x.shape[1]               # target rows = rows of u; x may hold more rows (sources that are not targets)
    nnz = max(col.numel(), 1)
    a_pre = torch.empty((nnz, H), dtype=torch.float32, device=x.device)
    a_post = torch.empty((nnz, H), dtype=torch.float32, device=x.device) if p_drop > 0 else None
    ssum = torch.empty((N, H), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().b2g_tz_alpha(_p(x), _ld(x), _p(u), _ld(u), N, H, F, _dt(x), _p(rowptr), _p(col), _p(a_pre), _p(a_post),
                                        _p(ssum), _p(edge_bias), float(p_drop), int(seed), int(band), int(impl), _stream()), "tz_alpha")
    return a_pre, a_post, ssum


def edge_rows_sl(edge_attr, eid, rowptr, n_rows):
    """GATConv(edge_dim): fp32 [nnz, 4] edge attributes in the order of the self-loop-replaced target-major CSR, the new
    self loops filled with the mean attribute of the row's other entries (PyG fill_value='mean')."""
    _cuda(edge_attr, eid, rowptr)
    ea = edge_attr.detach().float().contiguous()
    out = torch.empty((max(eid.numel(), 1), 4), dtype=torch.float32, device=eid.device)
    _lib.check(_lib.load().b2g_edge_rows_sl(_p(ea), ea.shape[0], _p(eid), _p(rowptr), n_rows, _p(out), _stream()), "edge_rows_sl")
    return out


def gatz_bwd(x, a, dz, g, H, slope, csr, csr_t, perm, smax, ssum, p_drop, seed, y_out, band=0, d_a_out=None, edge_bias=None,
             want_de=False, max_row_len=0):
    """Writes y = [sum_i alpha_ij1 g_i | ...] into y_out (a [N, H*C] view, any row stride) and the logit gradients
    d a = [d a_src | d a_dst] into d_a_out (a [N, 2H] view of dtype fp32 or x.dtype, any row stride; default: a new fp32
    tensor).  Returns d a."""
    lib = _lib.load()
    x, dz, g = _rows(x), _rows(dz), _rows(g)
    N, F = x.shape
    C = g.shape[1]
    dev = x.device
    nnz = csr[1].numel()
    alpha_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    de_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    d_a = d_a_out if d_a_out is not None else torch.empty((N, 2 * H), dtype=torch.float32, device=dev)
    assert d_a.shape == (N, 2 * H) and d_a.stride(1) == 1 and d_a.dtype in (torch.float32, torch.bfloat16)
    st = _stream()
    _lib.check(lib.b2g_gatz_bwd_dst(_p(x), _ld(x), _p(a), a.stride(0), _p(dz), _ld(dz), N, H, F, _dt(x), float(slope),
                                    _p(csr[0]), _p(csr[1]), _p(smax), _p(ssum), float(p_drop), int(seed), _p(alpha_e),
                                    _p(de_e), _p(d_a), _ld(d_a), _dt(d_a), _p(edge_bias), int(band), int(max_row_len), st),
               "gatz_bwd_dst")
    _lib.check(lib.b2g_gatz_bwd_src(_p(g), _ld(g), _p(alpha_e), _p(de_e), _p(y_out), _ld(y_out), _p(d_a), _ld(d_a), _dt(d_a),
                                    N, H, C, _dt(x), _p(csr_t[0]), _p(csr_t[1]), _p(perm), int(band), st), "gatz_bwd_src")
    return (d_a, de_e) if want_de else d_a                 # de_e: gradient of the logits in front of the LeakyReLU, [nnz, H]


def seg_wsum4(x, w_e, rowptr, col, perm, out, d_a=None, band=0):
    """out[i] = [sum_t w_e[p_t,0] x[col_t] | ... | sum_t w_e[p_t,3] x[col_t]] over row i of (rowptr, col), p_t = perm[t] or t
    (gat_rows.cu gatz_bwd_src_kernel).  d_a (optional, [N, >= 4] view, fp32 or bf16) receives the row sums of w_e's companion."""
    lib = _lib.load()
    x = _rows(x)
    N = out.shape[0]
    C = x.shape[1]
    _lib.check(lib.b2g_gatz_bwd_src(_p(x), _ld(x), _p(w_e), _p(w_e), _p(out), _ld(out), _p(d_a),
                                    d_a.stride(0) if d_a is not None else 0, _dt(d_a) if d_a is not None else 0, N, 4, C,
                                    _dt(x), _p(rowptr), _p(col), _p(perm), int(band), _stream()), "seg_wsum4")
    return out


def tz_fwd(x, u, H, rowptr, col, p_drop, seed, save_alpha, band=0, edge_bias=None, extra_cols=0, x_self=None, alpha_out=None):
    """z_aug [N, H*F + 8 + F (+ extra_cols, left for the caller to fill)] (see include/b2g.h b2g_tz_fwd) and the
    pre-dropout attention weights [nnz, H] | None.  edge_bias: fp32 [nnz, H] added to the logits (edge features)."""
    x, u = _rows(x), _rows(u)
    N, F = u.shape[0], x.shape[1]                      # rows of this call = rows of u (a row range when x_self is given)
    z = empty_rows(N, H * F + 8 + F + extra_cols, x.dtype, x.device)
    alpha = alpha_out if alpha_out is not None else (
        torch.empty((max(col.numel(), 1), H), dtype=torch.float32, device=x.device) if save_alpha else None)
    _lib.check(_lib.load().b2g_tz_fwd(_p(x), _ld(x), _p(x_self), _p(u), _ld(u), _p(z), _ld(z), N, H, F, _dt(x), _p(rowptr), _p(col),
                                      _p(alpha), _p(edge_bias), float(p_drop), int(seed), int(band), _stream()), "tz_fwd")
    return z, alpha


def tz_bwd_dst(x, dz_aug, alpha, H, rowptr, col, p_drop, seed, du_out, band=0, edge_bias=None, max_row_len=0):
    """-> (alpha_e after dropout, de_e) fp32 [nnz, H], target-major; writes du = [sum_j de_ijh x_j]_h into du_out
    (a [N, H*F] view, any row stride) from the same gather.  edge_bias: fp32 [nnz, H] added to d alpha' (edge features)."""
    x, dz_aug = _rows(x), _rows(dz_aug)
    N, F = x.shape
    alpha_e = torch.empty_like(alpha)
    de_e = torch.empty_like(alpha)
    _lib.check(_lib.load().b2g_tz_bwd_dst(_p(x), _ld(x), _p(dz_aug), _ld(dz_aug), _p(alpha), N, H, F, _dt(x), _p(rowptr),
                                          _p(col), float(p_drop), int(seed), _p(alpha_e), _p(de_e), _p(du_out),
                                          _ld(du_out) if du_out is not None else 0, _p(edge_bias), int(band), int(max_row_len),
                                          _stream()),
               "tz_bwd_dst")
    return alpha_e, de_e


def edge_dot4(v, ea_csr, rowptr, H):
    """out[p, h] = v[i(p), 4h..4h+3] . ea_csr[p]: v fp32 [N, >= 4H] (any row stride), ea_csr fp32 [nnz, 4] -> fp32 [nnz, H]."""
    _cuda(v, ea_csr, rowptr)
    assert v.dtype == torch.float32 and ea_csr.dtype == torch.float32 and v.stride(1) == 1 and ea_csr.is_contiguous()
    N = v.shape[0]                                          # v.stride(0) == 0 (an expanded [1, 4H] row): the same v for all nodes
    out = torch.empty((max(ea_csr.shape[0], 1), H), dtype=torch.float32, device=v.device)
    _lib.check(_lib.load().b2g_edge_dot4(_p(v), v.stride(0), _p(ea_csr), _p(rowptr), N, H, _p(out), _stream()), "edge_dot4")
    return out


def edge_wsum4(w, ea_csr, rowptr, H, out, p_drop=0.0, seed=0):
    """out[i, 4h + c] = sum_p w[p, h] keep(p, h) ea_csr[p, c] into `out` ([N, 4H] view of x's dtype, any row stride)."""
    _cuda(w, ea_csr, rowptr, out)
    assert w.dtype == torch.float32 and out.stride(1) == 1 and out.shape[1] == 4 * H
    _lib.check(_lib.load().b2g_edge_wsum4(_p(w), _p(ea_csr), _p(rowptr), out.shape[0], H, float(p_drop), int(seed), _p(out),
                                          out.stride(0), _dt(out), _stream()), "edge_wsum4")
    return out


# ------------------------------------------------------------------------------------------ K5
def tconv_fwd(q, k, v, skip, H, C, concat, rowptr, col, p_drop, seed, save_stats):
    lib = _lib.load()
    N = q.shape[0]
    out = torch.empty((N, H * C if concat else C), dtype=q.dtype, device=q.device)
    smax = torch.empty((N, H), dtype=torch.float32, device=q.device) if save_stats else None
    ssum = torch.empty((N, H), dtype=torch.float32, device=q.device) if save_stats else None
    assert q.stride(0) == k.stride(0) == v.stride(0)
    _lib.check(lib.b2g_tconv_fwd(_p(q), _p(k), _p(v), _ld(q), _p(skip), _ld(skip) if skip is not None else 0, _p(out),
                                 _ld(out), N, H, C, _dt(q), int(concat), _p(rowptr), _p(col), _p(smax), _p(ssum),
                                 float(p_drop), int(seed), _stream()), "tconv_fwd")
    return out, smax, ssum


def tconv_bwd(q, k, v, gout, H, C, concat, csr, csr_t, perm, smax, ssum, p_drop, seed, dq, dk, dv):
    """Writes dq, dk, dv (views [N,H*C] sharing one row stride)."""
    lib = _lib.load()
    N = q.shape[0]
    dev = q.device
    nnz = csr[1].numel()
    alpha_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    ds_e = torch.empty((max(nnz, 1), H), dtype=torch.float32, device=dev)
    gout = _rows(gout)
    st = _stream()
    assert dk.stride(0) == dv.stride(0)
    _lib.check(lib.b2g_tconv_bwd_dst(_p(q), _p(k), _p(v), _ld(q), _p(gout), _ld(gout), N, H, C, _dt(q), int(concat),
                                     _p(csr[0]), _p(csr[1]), _p(smax), _p(ssum), float(p_drop), int(seed),
                                     _p(alpha_e), _p(ds_e), _p(dq), _ld(dq), st), "tconv_bwd_dst")
    _lib.check(lib.b2g_tconv_bwd_src(_p(q), _ld(q), _p(gout), _ld(gout), _p(alpha_e), _p(ds_e), _p(dk), _p(dv),
                                     _ld(dk), N, H, C, _dt(q), int(concat), _p(csr_t[0]), _p(csr_t[1]), _p(perm),
                                     st), "tconv_bwd_src")


# ------------------------------------------------------------------------------------------ halo
def rows_gather(x, idx, out=None):
    """out[r] = x[idx[r]]; idx None: out[r] = x[r] for every row of x (a copy into `out`, e.g. a column block of a wider
    matrix: torch splits such a copy of > 2^31 elements into ~30 launches at a third of the bandwidth)."""
    lib = _lib.load()
    x = _rows(x)
    n = idx.numel() if idx is not None else x.shape[0]
    if out is None:
        out = torch.empty((n, x.shape[1]), dtype=x.dtype, device=x.device)
    _lib.check(lib.b2g_rows_gather(_p(x), _ld(x), _p(idx), n, _p(out), _ld(out), x.shape[1], _dt(x), _stream()), "rows_gather")
    return out


def rows_scatter_add(x, idx, src):
    lib = _lib.load()
    src = _rows(src)
    _lib.check(lib.b2g_rows_scatter_add(_p(x), _ld(x), _p(idx), idx.numel(), _p(src), _ld(src), x.shape[1], _dt(x),
                                        _stream()), "rows_scatter_add")
    return x


# ------------------------------------------------------------------------------------------ BatchNorm + fused glue
def bn_supported(x) -> bool:
    if not x.is_cuda or x.dim() != 2 or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    rb = x.shape[1] * x.element_size()
    return x.shape[0] >= 1 and rb % 16 == 0 and rb <= 4096


def bn_stats(x, r, eps: float) -> torch.Tensor:
    """[3, C] fp32: mean, 1/sqrt(var + eps), biased var of s = x (+ r) over the rows."""
    lib = _lib.load()
    x = _rows(x)
    r = _rows(r) if r is not None else None
    N, C = x.shape
    stats = torch.empty((3, C), dtype=torch.float32, device=x.device)
    ws = _ws(lib.b2g_bn_workspace_bytes(C), x.device)
    _lib.check(lib.b2g_bn_stats(_p(x), _ld(x), _p(r), _ld(r) if r is not None else 0, N, C, _dt(x), float(eps), _p(stats),
                                _p(ws), _stream()), "bn_stats")
    return stats


def bn_apply(x, r, mean, rstd, gamma, beta, relu: bool, p_drop: float, seed: int, keep_s: bool):
    """y = dropout(relu(gamma (s - mean) rstd + beta)), s = x (+ r).  Returns (y, s | None)."""
    lib = _lib.load()
    x = _rows(x)
    r = _rows(r) if r is not None else None
    N, C = x.shape
    y = torch.empty((N, C), dtype=x.dtype, device=x.device)
    s = torch.empty((N, C), dtype=x.dtype, device=x.device) if (keep_s and r is not None) else None
    _lib.check(lib.b2g_bn_apply(_p(x), _ld(x), _p(r), _ld(r) if r is not None else 0, _p(y), _ld(y), _p(s),
                                _ld(s) if s is not None else 0, N, C, _dt(x), _p(mean), _p(rstd), _p(gamma), _p(beta),
                                int(relu), float(p_drop), int(seed), _stream()), "bn_apply")
    return y, s


def bn_bwd(dy, y, s, mean, rstd, gamma, relu: bool, drop_scale: float, training: bool, reduce_sums=None):
    """-> (ds [N,C], sums fp32 [2,C] = (sum dz, sum dz * xhat) over THIS process's rows).
    `reduce_sums(sums, n_local) -> sums'` (optional) maps the local column sums to what the elementwise pass should use
    (multi-GPU BatchNorm: all-reduce, rescaled so that sums' / n_local == global sums / global N)."""
    lib = _lib.load()
    dy, s = _rows(dy), _rows(s)
    y = _rows(y) if y is not None else None
    N, C = s.shape
    sums = torch.empty((2, C), dtype=torch.float32, device=s.device)
    ws = _ws(lib.b2g_bn_workspace_bytes(C), s.device)
    st = _stream()
    _lib.check(lib.b2g_bn_bwd_stats(_p(dy), _ld(dy), _p(y), _ld(y) if y is not None else 0, _p(s), _ld(s), N, C, _dt(s),
                                    _p(mean), _p(rstd), int(relu), float(drop_scale), _p(sums), _p(ws), st), "bn_bwd_stats")
    used = reduce_sums(sums, N).contiguous() if reduce_sums is not None else sums
    ds = torch.empty((N, C), dtype=s.dtype, device=s.device)
    _lib.check(lib.b2g_bn_bwd_apply(_p(dy), _ld(dy), _p(y), _ld(y) if y is not None else 0, _p(s), _ld(s), _p(ds), _ld(ds),
                                    N, C, _dt(s), _p(mean), _p(rstd), _p(gamma), _p(used), int(relu), float(drop_scale),
                                    int(training), _stream()), "bn_bwd_apply")
    return ds, sums
