"""torch.autograd.Function wrappers over the C ABI (include/eqv2_b200.h).

Each Function's forward/backward launches hand-written kernels on the current CUDA stream;
nothing here falls back to PyTorch math for the operators of SURVEY §8a.  Host-side tables
(coefficient layouts, CSR edge plans, grid matrices) are built once and cached.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, _so3_math

_F32 = torch.float32


# ----------------------------------------------------------------------------------------------
# coefficient layout tables  (so3.py:45-199 CoefficientMappingModule, single resolution)
# ----------------------------------------------------------------------------------------------
class CoeffLayout:
    _cache = {}

    def __init__(self, lmax, mmax):
        self.lmax, self.mmax = lmax, mmax
        self.K = (lmax + 1) ** 2
        self.WS = (lmax + 1) * (4 * (lmax + 1) ** 2 - 1) // 3
        red = [(l, m) for l in range(lmax + 1) for m in range(-min(l, mmax), min(l, mmax) + 1)]
        self.Kr = len(red)
        self.full_of_red = [l * l + l + m for l, m in red]               # l-primary reduced -> full index
        # m-primary order: m=0 rows (l=0..L), then for m=1..M: +m rows, -m rows
        order = [(l, 0) for l in range(lmax + 1)]
        self.m_sizes = [lmax + 1]
        for m in range(1, mmax + 1):
            order += [(l, m) for l in range(m, lmax + 1)]
            order += [(l, -m) for l in range(m, lmax + 1)]
            self.m_sizes.append(lmax - m + 1)
        self.m_order = order
        red_index = {lm: i for i, lm in enumerate(red)}
        self.to_m = [red_index[lm] for lm in order]                      # m_primary[j] = l_primary[to_m[j]]
        pos_of_full = [-1] * self.K
        for p, (l, m) in enumerate(order):
            pos_of_full[l * l + l + m] = p
        self.pos_of_full_host = pos_of_full
        # radial slot per m-primary row: slot * Cin + channel indexes the rad vector (so2_ops.py:100-133)
        slot, base = [], 0
        for m in range(mmax + 1):
            n = lmax - m + 1
            if m == 0:
                slot += [base + i for i in range(n)]
            else:
                slot += [base + i for i in range(n)] * 2
            base += n
        self.rad_slot_host = slot
        self.nslot = base
        self.row_l = [l for l, _ in order]
        self._dev = {}

    @classmethod
    def get(cls, lmax, mmax):
        key = (lmax, mmax)
        if key not in cls._cache:
            cls._cache[key] = cls(lmax, mmax)
        return cls._cache[key]

    def dev(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = dict(
                pos_of_full=torch.tensor(self.pos_of_full_host, dtype=torch.int32, device=device),
                rad_slot=torch.tensor(self.rad_slot_host, dtype=torch.int32, device=device),
                to_m=torch.tensor(self.to_m, dtype=torch.long, device=device),
                jd=torch.from_numpy(_so3_math.jd_packed(self.lmax)).to(device),
            )
        return self._dev[key]

    def conv_groups(self, c_in, c_out, extra):
        """(a_off, k_g, c_off, n_g) of the m = 0..mmax blocks of one SO(2) convolution, operands in
        m-primary order: A[E, Kr*c_in] -> Y[E, extra + Kr*c_out]."""
        groups = []
        a_off, c_off = 0, 0
        for m, n in enumerate(self.m_sizes):
            rows = n if m == 0 else 2 * n
            k_g = rows * c_in
            n_g = rows * c_out + (extra if m == 0 else 0)
            groups.append((a_off, k_g, c_off, n_g))
            a_off += k_g
            c_off += n_g
        return groups


# ----------------------------------------------------------------------------------------------
# edge plan: dst-sorted / src-sorted CSR over the (arbitrarily ordered) edge list
# ----------------------------------------------------------------------------------------------
def _csr_from_index(idx, num_nodes):
    """(perm int32 [E], rowptr int32 [N+1]): edges grouped by idx, stable in edge order
    (`eqv2_csr_from_index`: histogram -> scan -> fill -> per-bucket sort)."""
    dev = idx.device
    E = int(idx.shape[0])
    counts = torch.zeros(num_nodes, dtype=torch.int32, device=dev)
    cursor = torch.zeros(num_nodes, dtype=torch.int32, device=dev)
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev)
    _lib.call("eqv2_csr_from_index", idx.data_ptr(), E, num_nodes, counts.data_ptr(), rowptr.data_ptr(),
              cursor.data_ptr(), perm.data_ptr(), _lib.stream_ptr(), n_kernels=4)
    return perm, rowptr


def _csr_few_buckets(idx, num_buckets):
    """(perm, rowptr) like `_csr_from_index` for FEW, LARGE buckets (element types: thousands of edges each; the
    per-bucket ordering pass of eqv2_csr_from_index is sized for node degrees).  Once per graph: a stable sort."""
    perm = torch.argsort(idx, stable=True).to(torch.int32)
    counts = torch.zeros(num_buckets + 1, dtype=torch.int32, device=idx.device)
    counts.scatter_add_(0, idx + 1, torch.ones_like(idx, dtype=torch.int32))     # (torch.bincount reads back the max)
    return perm, torch.cumsum(counts, 0, dtype=torch.int32)


class EdgePlan:
    """CSR views of one edge list.  `rowptr_dst`/`perm_dst` group edges by destination (segment softmax,
    segmented reduce), `rowptr_src`/`perm_src` by source (node-centric backward of the gather)."""

    def __init__(self, edge_index, num_nodes, dst_sorted_rowptr=None):
        _lib.check_device(edge_index)
        self.E = int(edge_index.shape[1])
        self.N = int(num_nodes)
        self.src = edge_index[0].contiguous()
        self.dst = edge_index[1].contiguous()
        dev = edge_index.device
        if dst_sorted_rowptr is not None:
            # the neighbour-list builders emit edges already sorted by destination
            self.rowptr_dst = dst_sorted_rowptr
            self.perm_dst = torch.arange(self.E, dtype=torch.int32, device=dev)
        else:
            self.perm_dst, self.rowptr_dst = _csr_from_index(self.dst, self.N)
        self.perm_src, self.rowptr_src = _csr_from_index(self.src, self.N)
        self._z = None

    def element_types(self, atomic_numbers, num_elements):
        """(Z[src], csr, Z[dst], csr): per-edge element types and their CSR over the embedding rows, shared by every
        block of a forward pass (the reference recomputes atomic_numbers[edge_index[i]] per block)."""
        key = (atomic_numbers.data_ptr(), atomic_numbers._version, int(num_elements))
        if self._z is None or self._z[0] != key:
            zs, zd = atomic_numbers[self.src].contiguous(), atomic_numbers[self.dst].contiguous()
            self._z = (key, (zs, _csr_few_buckets(zs, num_elements), zd, _csr_few_buckets(zd, num_elements)), atomic_numbers)
        return self._z[1]


_plan_cache = {}


def _plan_key(edge_index, num_nodes):
    return (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), str(edge_index.device))


def edge_plan(edge_index, num_nodes):
    key = _plan_key(edge_index, num_nodes)
    hit = _plan_cache.get("k")
    if hit is not None and hit[0] == key:
        return hit[1]
    plan = EdgePlan(edge_index, num_nodes)
    _plan_cache["k"] = (key, plan, edge_index)   # keep the tensor alive so data_ptr stays unique
    return plan


def _register_plan(edge_index, num_nodes, rowptr_dst):
    plan = EdgePlan(edge_index, num_nodes, dst_sorted_rowptr=rowptr_dst)
    _plan_cache["k"] = (_plan_key(edge_index, num_nodes), plan, edge_index)
    return plan


# ----------------------------------------------------------------------------------------------
# neighbour lists
# ----------------------------------------------------------------------------------------------
def _graph_ptr(natoms, dev):
    B = int(natoms.shape[0])
    natoms = natoms.to(device=dev, dtype=torch.long).contiguous()
    gp = torch.empty(B + 1, dtype=torch.int32, device=dev)
    _lib.call("eqv2_graph_ptr", natoms.data_ptr(), gp.data_ptr(), B, _lib.stream_ptr())
    return gp


def _count_scan_fill(N, dev, launch):
    """count (mode 0) -> exclusive scan -> read E -> fill (mode 1).  `launch(mode, deg, rowptr, outs)`."""
    deg = torch.empty(N, dtype=torch.int32, device=dev)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    launch(0, deg, rowptr, None, err)
    _lib.call("eqv2_exclusive_scan", deg.data_ptr(), rowptr.data_ptr(), N, _lib.stream_ptr())
    tail = torch.stack([rowptr[-1], err[0]]).tolist()          # the one host read-back of the builder
    E = int(tail[0])
    if tail[1] != 0:
        raise _lib.Eqv2Error("neighbour list: a destination atom has more in-cutoff candidates than the kernel holds")
    edge_index = torch.empty(2, E, dtype=torch.long, device=dev)
    dist = torch.empty(E, dtype=_F32, device=dev)
    vec = torch.empty(E, 3, dtype=_F32, device=dev)
    if E > 0:
        launch(1, deg, rowptr, (edge_index, dist, vec), err)
    return edge_index, dist, vec, rowptr


def radius_graph(pos, natoms, batch, cutoff, max_neighbors):
    """Isolated-molecule radius graph (equiformerv2_qm9.py:423-525) -> (edge_index [2,E] int64 with
    row 0 = source, row 1 = destination; edge_distance [E]; edge_vec [E,3] = pos[dst] - pos[src]),
    edges sorted by (dst, src).  No gradient (the reference QM9/OC20 models never differentiate it)."""
    _lib.check_device(pos, batch)
    pos = pos.detach().to(_F32).contiguous()
    batch = batch.contiguous()
    N = int(pos.shape[0])
    dev = pos.device
    gp = _graph_ptr(natoms, dev)
    mx = -1 if max_neighbors is None else int(max_neighbors)

    def launch(mode, deg, rowptr, outs, err):
        ei, d, v = outs if outs is not None else (None, None, None)
        _lib.call("eqv2_radius_graph", pos.data_ptr(), gp.data_ptr(), batch.data_ptr(), N, float(cutoff), mx, mode,
                  deg.data_ptr(), rowptr.data_ptr(), _lib.ptr(ei[0]) if ei is not None else None,
                  _lib.ptr(ei[1]) if ei is not None else None, _lib.ptr(d), _lib.ptr(v), err.data_ptr(),
                  _lib.stream_ptr())

    ei, d, v, rowptr = _count_scan_fill(N, dev, launch)
    _register_plan(ei, N, rowptr)
    return ei, d, v


def radius_graph_pbc(pos, cell, natoms, batch, cutoff, max_neighbors, strict=False):
    """Periodic radius graph with the semantics of the fairchem `generate_graph` call at
    equiformerv2_oc20.py:223-234 as restated in oracle/eqv2_oracle.py::radius_graph_pbc_fairchem
    (row 0 = neighbour j, row 1 = centre i, vec = pos[j] + offset - pos[i]); edges sorted by (i, j, image)."""
    _lib.check_device(pos, cell, batch)
    pos = pos.detach().to(_F32).contiguous()
    cell = cell.detach().to(_F32).contiguous()
    batch = batch.contiguous()
    N, B = int(pos.shape[0]), int(cell.shape[0])
    dev = pos.device
    gp = _graph_ptr(natoms, dev)
    reps = torch.empty(B, 3, dtype=torch.int32, device=dev)
    _lib.call("eqv2_pbc_reps", cell.data_ptr(), float(cutoff), reps.data_ptr(), B, _lib.stream_ptr())

    def launch(mode, deg, rowptr, outs, err):
        ei, d, v = outs if outs is not None else (None, None, None)
        _lib.call("eqv2_radius_graph_pbc", pos.data_ptr(), cell.data_ptr(), gp.data_ptr(), batch.data_ptr(),
                  reps.data_ptr(), N, float(cutoff), int(max_neighbors), int(bool(strict)), mode, deg.data_ptr(),
                  rowptr.data_ptr(), _lib.ptr(ei[0]) if ei is not None else None,
                  _lib.ptr(ei[1]) if ei is not None else None, _lib.ptr(d), _lib.ptr(v), err.data_ptr(),
                  _lib.stream_ptr())

    ei, d, v, rowptr = _count_scan_fill(N, dev, launch)
    _register_plan(ei, N, rowptr)
    return ei, d, v


def radius_graph_matpes(pos, cell, natoms, batch, cutoff, max_neighbors, version):
    """MatPES periodic builders (27 images).  version 1: equiformerv2_MatPES.py:258-340; version 2:
    equiformerv2_MatPESv2.py:177-240.  Returns (edge_index [2,E] row 0 = src, row 1 = dst, edge_distance,
    edge_vec, image index int32 [E]); topology only -- no gradient (the reference builds it on detached
    positions in v2; differentiable vectors are rebuilt from pos/cell by the model wrapper)."""
    _lib.check_device(pos, cell, batch)
    pos = pos.detach().to(_F32).contiguous()
    cell = cell.detach().to(_F32).contiguous()
    batch = batch.contiguous()
    N = int(pos.shape[0])
    dev = pos.device
    gp = _graph_ptr(natoms, dev)
    mx = -1 if max_neighbors is None else int(max_neighbors)
    holder = {}

    def launch(mode, deg, rowptr, outs, err):
        ei, d, v = outs if outs is not None else (None, None, None)
        img = torch.empty(ei.shape[1], dtype=torch.int32, device=dev) if ei is not None else None
        holder["img"] = img
        _lib.call("eqv2_radius_graph_pbc27", pos.data_ptr(), cell.data_ptr(), gp.data_ptr(), batch.data_ptr(), N,
                  float(cutoff), mx, int(version), mode, deg.data_ptr(), rowptr.data_ptr(),
                  _lib.ptr(ei[0]) if ei is not None else None, _lib.ptr(ei[1]) if ei is not None else None,
                  _lib.ptr(img), _lib.ptr(d), _lib.ptr(v), err.data_ptr(), _lib.stream_ptr())

    ei, d, v, rowptr = _count_scan_fill(N, dev, launch)
    img = holder.get("img")
    if img is None:
        img = torch.empty(0, dtype=torch.int32, device=dev)
    _register_plan(ei, N, rowptr)
    return ei, d, v, img


class SegmentSumFn(torch.autograd.Function):
    """Per-graph sum of per-atom scalars; `batch` must be non-decreasing (every reference collate
    function produces it that way).  Deterministic replacement of the index_add_ readouts."""

    @staticmethod
    def forward(ctx, values, batch, num_graphs):
        _lib.check_device(values, batch)
        assert values.is_contiguous() and batch.is_contiguous()
        N = int(values.shape[0])
        out = torch.empty(num_graphs, dtype=_F32, device=values.device)
        _lib.call("eqv2_segment_sum_fwd", values.data_ptr(), 1, batch.data_ptr(), out.data_ptr(), N, int(num_graphs),
                  _lib.stream_ptr())
        ctx.save_for_backward(batch)
        ctx.num_graphs = int(num_graphs)
        return out

    @staticmethod
    def backward(ctx, gout):
        (batch,) = ctx.saved_tensors
        return SegmentBcastFn.apply(gout.contiguous(), batch, ctx.num_graphs), None, None


class SegmentBcastFn(torch.autograd.Function):
    """gv[n] = gout[batch[n]] -- the adjoint of SegmentSumFn (and vice versa)."""

    @staticmethod
    def forward(ctx, gout, batch, num_graphs):
        assert gout.is_contiguous()
        N = int(batch.shape[0])
        gv = torch.empty(N, dtype=_F32, device=gout.device)
        _lib.call("eqv2_segment_sum_bwd", gout.data_ptr(), batch.data_ptr(), gv.data_ptr(), N, _lib.stream_ptr())
        ctx.save_for_backward(batch)
        ctx.num_graphs = num_graphs
        return gv

    @staticmethod
    def backward(ctx, ggv):
        (batch,) = ctx.saved_tensors
        return SegmentSumFn.apply(ggv.contiguous(), batch, ctx.num_graphs), None, None


def segment_sum_nodes(values, batch, num_graphs):
    return SegmentSumFn.apply(values.contiguous(), batch.contiguous(), num_graphs)


# ----------------------------------------------------------------------------------------------
# row gathers / deterministic segmented column sums (embedding lookups and their gradients, bias gradients)
# ----------------------------------------------------------------------------------------------
def _seg_colsum(src, ld, col_off, rows, C, V, rowptr, perm, S):
    partial = torch.empty(V * S * C, dtype=_F32, device=src.device)
    out = torch.empty(V, C, dtype=_F32, device=src.device)
    _lib.call("eqv2_seg_colsum", src.data_ptr() + 4 * col_off, int(ld), _lib.ptr(rowptr), _lib.ptr(perm), int(rows),
              int(V), int(C), int(S), partial.data_ptr(), out.data_ptr(), _lib.stream_ptr(), n_kernels=2,
              work=(0.0, 4.0 * rows * C))
    return out


class EmbedRowsFn(torch.autograd.Function):
    """out[e] = table[idx[e]]; `csr` = (perm, rowptr) of idx over the table rows (for the deterministic backward)."""

    @staticmethod
    def forward(ctx, table, idx, csr):
        _lib.check_device(table, idx)
        assert table.is_contiguous() and idx.is_contiguous() and idx.dtype == torch.long
        E, C = int(idx.shape[0]), int(table.shape[1])
        out = torch.empty(E, C, dtype=_F32, device=table.device)
        _lib.call("eqv2_embed_rows", table.data_ptr(), idx.data_ptr(), out.data_ptr(), E, C, _lib.stream_ptr(),
                  work=(0.0, 8.0 * E * C))
        ctx.save_for_backward(idx)
        ctx.csr, ctx.V = csr, int(table.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return EmbedRowsBwdFn.apply(g.contiguous(), idx, ctx.csr, ctx.V), None, None


class EmbedRowsBwdFn(torch.autograd.Function):
    """gtable[v] = sum_{e: idx[e] = v} g[e] in a fixed order -- the adjoint of EmbedRowsFn (and vice versa)."""

    @staticmethod
    def forward(ctx, g, idx, csr, V):
        _lib.check_device(g, idx)
        perm, rowptr = csr
        ctx.save_for_backward(idx)
        ctx.csr = csr
        # few element types occur in a batch -> few non-empty segments: split each into many row groups
        S = max(1, min(256, int(g.shape[0]) // 32))
        return _seg_colsum(g, g.shape[1], 0, g.shape[0], g.shape[1], V, rowptr, perm, S)

    @staticmethod
    def backward(ctx, gg):
        (idx,) = ctx.saved_tensors
        return EmbedRowsFn.apply(gg.contiguous(), idx, ctx.csr), None, None, None


def embed_rows(table, idx, csr):
    return EmbedRowsFn.apply(table.contiguous(), idx.contiguous(), csr)


class BlockWeightFn(torch.autograd.Function):
    """B[2h, 2k] = [[Wr, -Wi], [Wi, Wr]] from W[2h, k] = [Wr; Wi]  (so2_ops.py:53-61 folded into the GEMM operand)."""

    @staticmethod
    def forward(ctx, W):
        _lib.check_device(W)
        assert W.is_contiguous() and W.dim() == 2 and W.shape[0] % 2 == 0
        h, k = int(W.shape[0]) // 2, int(W.shape[1])
        B = torch.empty(2 * h, 2 * k, dtype=_F32, device=W.device)
        _lib.call("eqv2_so2_block_weight", W.data_ptr(), B.data_ptr(), h, k, _lib.stream_ptr())
        return B

    @staticmethod
    def backward(ctx, gB):
        return BlockWeightAdjFn.apply(gB.contiguous())


class BlockWeightAdjFn(torch.autograd.Function):
    """gW[2h, k] = [gB00 + gB11; gB10 - gB01] -- the adjoint of BlockWeightFn (and vice versa: both maps are linear)."""

    @staticmethod
    def forward(ctx, gB):
        _lib.check_device(gB)
        assert gB.is_contiguous()
        h, k = int(gB.shape[0]) // 2, int(gB.shape[1]) // 2
        gW = torch.empty(2 * h, k, dtype=_F32, device=gB.device)
        _lib.call("eqv2_so2_block_weight_adj", gB.data_ptr(), gW.data_ptr(), h, k, _lib.stream_ptr())
        return gW

    @staticmethod
    def backward(ctx, ggW):
        return BlockWeightFn.apply(ggW.contiguous())


def so2_block_weight(W):
    return BlockWeightFn.apply(W.contiguous())


class ColsumFn(torch.autograd.Function):
    """X[:, off:off+n].sum(0) for a contiguous matrix, fixed summation order (bias gradients)."""

    @staticmethod
    def forward(ctx, X, off, n):
        _lib.check_device(X)
        assert X.is_contiguous() and X.dim() == 2
        ctx.spec = (tuple(X.shape), off, n)
        rows = int(X.shape[0])
        if rows == 0:
            return torch.zeros(n, dtype=_F32, device=X.device)
        S = max(1, min(rows // 16, -(-2368 // ((n + 127) // 128))))      # ~16 CTAs of 128 threads per SM
        return _seg_colsum(X, X.shape[1], off, rows, n, 1, None, None, S).view(n)

    @staticmethod
    def backward(ctx, g):
        shape, off, n = ctx.spec
        gX = g.new_zeros(shape)
        gX[:, off:off + n] = g
        return gX, None, None


def colsum(X, off, n):
    return ColsumFn.apply(X, off, n)


# ----------------------------------------------------------------------------------------------
# GEMM engine
# ----------------------------------------------------------------------------------------------
import os as _os

# A/B switches for measurements (bench.py --disable ...): "planes" = producer kernels write operand planes,
# "c_absmax" = the GEMM epilogue reduces max |C|.  Both on by default; EQV2_DISABLE=planes,c_absmax turns them off.
_FEATURES = {k: k not in _os.environ.get("EQV2_DISABLE", "").split(",") for k in ("planes", "c_absmax", "s2_planes")}

DEFAULT_GEMM_MODE = "f16x3"
_GEMM_MODE = {"mode": DEFAULT_GEMM_MODE}


def set_gemm_mode(mode):
    """Engine for the dense edge-level contractions:
      'f16x3' (default) : tcgen05 kind::f16 on operands pre-split into scaled fp16 hi/lo planes (three passes at
                          twice the TF32 rate, TMA operands, persistent CTAs) -- fp32-class accuracy,
      'tf32x3'          : tcgen05 kind::tf32 with the 3xTF32 hi/lo split done in the kernel -- fp32-class accuracy,
      'tf32'            : tcgen05, single TF32 pass (separately stated tolerance),
      'fp32'            : exact FFMA engine for everything.
      'f16'             : the f16x3 engine reading the hi planes only -- ONE fp16 tensor-core pass (reduced precision,
                          separately stated tolerance; BASELINE configs[3] "bf16/TF32 GEMM mode").
    Problems a tensor-core engine cannot address (two-level strided degree slabs, unaligned operands) run on the
    next engine down the list."""
    assert mode in ("fp32", "tf32x3", "tf32", "f16x3", "f16")
    _GEMM_MODE["mode"] = mode


def gemm_mode():
    return _GEMM_MODE["mode"]


class PlaneRef:
    """Stand-in for an fp32 matrix [rows, cols] that exists ONLY as the scaled fp16 hi/lo planes a producer kernel wrote
    (eqv2_*_planes entry points): it gives the GEMM descriptors an identity to find those planes by (split_scope) and the
    few tensor attributes the descriptor code looks at.  A launch with such an operand always runs on the f16 engine."""
    is_cuda = True
    dtype = _F32
    _version = 0

    def __init__(self, rows, cols, device):
        self.shape = (int(rows), int(cols))
        self.device = device

    def is_contiguous(self):
        return True

    def dim(self):
        return 2

    def data_ptr(self):
        return 0


class OperandSrc:
    """Where a GEMM operand lives inside the f16x3 split of tensor `t` (python side only): `t` read as a
    [rows, cols] matrix (slab_k > 0: a node tensor [n, slab_k, cols] repacked slab by slab), the operand's block
    starting at (row_off, col_off)."""
    __slots__ = ("t", "rows", "cols", "slab_k", "row_off", "col_off")

    def __init__(self, t, rows, cols, slab_k=0, row_off=0, col_off=0):
        self.t, self.rows, self.cols, self.slab_k = t, int(rows), int(cols), int(slab_k)
        self.row_off, self.col_off = int(row_off), int(col_off)

    @property
    def key(self):
        return (id(self.t), self.slab_k)


def _desc(A, B, C, bias, M, N, K, transA, transB, a_addr, b_addr, c_addr, a_off=0, b_off=0, c_off=0, accumulate=0,
          a_src=None, b_src=None):
    d = _lib.GemmDesc()
    d.A = A.data_ptr() + 4 * a_off
    d.B = B.data_ptr() + 4 * b_off
    d.C = C.data_ptr() + 4 * c_off
    d.bias = bias.data_ptr() if bias is not None else None
    d.M, d.N, d.K = int(M), int(N), int(K)
    d.transA, d.transB = int(transA), int(transB)
    d.a_rpb, d.a_bs, d.a_ld = a_addr
    d.b_rpb, d.b_bs, d.b_ld = b_addr
    d.c_rpb, d.c_bs, d.c_ld = c_addr
    d.accumulate = accumulate
    # python-side only: the tensors behind the raw pointers (plain matrices: the offset is a column offset)
    if a_src is None and A.dim() == 2:
        a_src = OperandSrc(A, A.shape[0], A.shape[1], 0, a_off // max(A.shape[1], 1), a_off % max(A.shape[1], 1))
    if b_src is None and B.dim() == 2:
        b_src = OperandSrc(B, B.shape[0], B.shape[1], 0, b_off // max(B.shape[1], 1), b_off % max(B.shape[1], 1))
    d.src = (a_src, b_src)
    return d


_BIG = 1 << 40


def _plain(ld):
    return (_BIG, 0, int(ld))


def _pick_split(descs, reduce_dim_large):
    """Split-K factor for reduction-heavy problems (weight gradients: K = #edges).
    In-kernel-split engines: the factor that fills whole waves of 148 CTAs best, >= 32 k-blocks (of 32) per split, <= 8.
    Persistent f16x3 engine: minimise  waves x (k-blocks per split + fixed per-tile cost)  with the measured per-tile cost
    of ~4 k-block times (promotion hand-over + epilogue, scripts/profile_step.py) + 2 for the atomic epilogue; up to 32
    splits, so a single 128 x 128 weight-gradient tile over 13 120 edges still spreads over 32 SMs."""
    if not reduce_dim_large:
        return 1
    tiles = sum(((d.M + 127) // 128) * ((d.N + 127) // 128) for d in descs)
    kmax = max(d.K for d in descs)
    if _GEMM_MODE["mode"] in ("f16x3", "f16") and _f16_ok(descs):
        nkb = (kmax + 63) // 64
        best, best_cost = 1, None
        for s in range(1, 33):
            if s > 1 and nkb // s < 4:
                break
            cost = ((tiles * s + 147) // 148) * ((nkb + s - 1) // s + (6.0 if s > 1 else 4.0))
            if best_cost is None or cost < best_cost * 0.97:
                best, best_cost = s, cost
        return best
    best, best_eff = 1, 0.0
    for s in range(1, 9):
        if s > 1 and kmax // (32 * s) < 32:
            break
        ctas = tiles * s
        eff = ctas / (148.0 * ((ctas + 147) // 148))
        if eff > best_eff + 0.04:
            best, best_eff = s, eff
    return best


def _tc_addressable(d):
    """Can the tensor-core engine address this problem? (see eqv2_gemm_tc in include/eqv2_b200.h)"""
    if d.a_ld % 4 or d.b_ld % 4 or (d.A % 16) or (d.B % 16):
        return False
    if (d.a_rpb < (1 << 31) and d.a_bs % 4) or (d.b_rpb < (1 << 31) and d.b_bs % 4):
        return False
    if (not d.transA and d.K % 4) or (d.transB and d.K % 4):
        return False
    return True


def _tc_ok(d):
    """Engine choice (measured, profiles/r01_bench_e): one CTA per 128x128 tile pays a fixed TMEM-allocation /
    pipeline-fill / epilogue cost, so short reductions (K < 256, i.e. < 8 k-blocks) with few output columns
    run faster on the FFMA engine."""
    if not _tc_addressable(d) or d.M * d.N * d.K < (1 << 21):
        return False
    return d.K >= 256 or d.N >= 1024


# ---- f16x3 engine: operand splits --------------------------------------------------------------------------
ABSMAX_SLOTS = 64        # EQV2_ABSMAX_SLOTS of csrc/common.cuh: a tensor's max |v| is the maximum over this many floats


class SplitF16:
    """Scaled fp16 hi/lo planes [2, rows, cols_pad] of a contiguous fp32 matrix + its absolute maximum (device)."""
    __slots__ = ("buf", "absmax", "rows", "cols", "cols_pad", "slab_k", "version")

    def __init__(self, src):
        self.rows, self.cols, self.slab_k = src.rows, src.cols, src.slab_k
        self.cols_pad = (self.cols + 63) // 64 * 64
        self.buf = torch.empty(2, self.rows, self.cols_pad, dtype=torch.float16, device=src.t.device)
        self.absmax = _zeroed_slot(src.t.device)
        self.version = src.t._version

    @property
    def plane(self):
        return self.rows * self.cols_pad


_ABSMAX = {}             # storage pointer -> (weakref(tensor), slot, version): max |v| written by the producing kernel


_SLOT_ARENA = {"buf": None, "next": 0}
_SLOT_ARENA_SIZE = 512


def _zeroed_slot(device):
    """One zero-initialised absmax slot (ABSMAX_SLOTS floats) out of an arena that is zero-filled 512 slots at a time:
    one fill launch per ~1.7 train steps instead of one fill / memset node per producer and per operand split."""
    a = _SLOT_ARENA
    if a["buf"] is None or a["next"] >= _SLOT_ARENA_SIZE or a["buf"].device != device:
        a["buf"] = torch.zeros(_SLOT_ARENA_SIZE, ABSMAX_SLOTS, dtype=_F32, device=device)
        a["next"] = 0
    slot = a["buf"][a["next"]]
    a["next"] += 1
    return slot


_ZERO_ARENA = {"buf": None, "next": 0}
_ZERO_ARENA_FLOATS = 1 << 16


def _small_zeros(shape, device):
    """Zero-initialised fp32 tensor for the small atomically-accumulated parameter gradients (LayerNorm / norm weights,
    attention dot vectors): carved out of a 256 KB arena that is zero-filled once per refill instead of one fill launch
    per tensor (tiny nodes are disproportionately expensive inside a replayed CUDA graph)."""
    n = 1
    for d in shape:
        n *= int(d)
    n4 = (n + 3) // 4 * 4
    if n4 > _ZERO_ARENA_FLOATS // 8:
        return torch.zeros(shape, dtype=_F32, device=device)
    a = _ZERO_ARENA
    if a["buf"] is None or a["next"] + n4 > _ZERO_ARENA_FLOATS or a["buf"].device != device:
        a["buf"] = torch.zeros(_ZERO_ARENA_FLOATS, dtype=_F32, device=device)
        a["next"] = 0
    out = a["buf"][a["next"]:a["next"] + n].view(shape)
    a["next"] += n4
    return out


def _absmax_slot(device):
    """Zero-initialised slot for a producer kernel's `absmax` output (only allocated in f16x3 mode)."""
    return _zeroed_slot(device) if _GEMM_MODE["mode"] in ("f16x3", "f16") else None


def _register_absmax(t, slot):
    """Remember that `slot` holds max |t| (the producing kernel reduced it while writing t)."""
    if slot is None:
        return
    import weakref
    key = t.untyped_storage().data_ptr()
    _ABSMAX[key] = (weakref.ref(t, lambda _r, k=key: _ABSMAX.pop(k, None) if _ABSMAX.get(k, (None,))[0] is _r else None),
                    slot, t._version)


def _known_absmax(t):
    """Slot with max |t| if t (or a same-size view of it) was written by an instrumented kernel and not modified since."""
    hit = _ABSMAX.get(t.untyped_storage().data_ptr())
    if hit is None:
        return None
    orig = hit[0]()
    if orig is None or orig._version != hit[2] or t._version != hit[2] or t.numel() != orig.numel() or t.storage_offset() != 0:
        return None
    return hit[1]


_SPLIT_SCOPES = []        # stack of {(id(tensor), slab_k): (tensor, SplitF16)}: splits valid while a backward pass runs
# NOTE: there is deliberately no cross-call cache of parameter splits.  `Tensor._version` is not a safe validity tag
# for parameters: the fused optimizers (torch.optim.AdamW(fused=True)) update them without bumping it.  A weight is
# split in the forward pass (~1 GB of traffic per step for 82.5 M parameters) and that split is reused by its dgrad.


class split_scope:
    """While active, GEMM calls find the splits in `pairs` [(tensor, SplitF16 | None)] by tensor identity and add the
    splits they make themselves (an output gradient is split once for its dgrad and its wgrad product)."""

    def __init__(self, pairs):
        self.d = {(id(t), sp.slab_k): (t, sp) for t, sp in pairs if sp is not None and t is not None}

    def __enter__(self):
        _SPLIT_SCOPES.append(self.d)
        return self

    def __exit__(self, *exc):
        _SPLIT_SCOPES.pop()
        return False


def reset_caches():
    """Drop every identity-keyed host cache (edge plans, producer maxima).  Called around CUDA-graph
    capture (graphs.py): work skipped because of a cache hit would be missing from the captured step."""
    _plan_cache.clear()
    _ABSMAX.clear()
    _SLOT_ARENA["buf"] = None          # slots handed out inside a capture must be zero-filled inside it
    _ZERO_ARENA["buf"] = None


def _find_split(src):
    t = src.t
    for d in reversed(_SPLIT_SCOPES):
        hit = d.get(src.key)
        if hit is not None and hit[0] is t and hit[1].version == t._version:
            return hit[1]
    return None


_SPLIT_LOG = None         # diagnostics: set to a list to record (rows, cols, maximum known) of every operand split


def _splits_for(srcs):
    """-> {src.key: SplitF16}; the missing splits are made by ONE eqv2_split_f16 call (two kernels)."""
    out, missing = {}, []
    for src in srcs:
        if src.key in out:
            continue
        sp = _find_split(src)
        if sp is None:
            sp = SplitF16(src)
            known = _known_absmax(src.t)
            if known is not None:
                sp.absmax = known
            missing.append((src, sp, known is not None))
        out[src.key] = sp
    for i in range(0, len(missing), _lib.MAX_SPLIT_ITEMS):
        part = missing[i:i + _lib.MAX_SPLIT_ITEMS]
        arr = (_lib.SplitDesc * len(part))()
        for a, (src, sp, given) in zip(arr, part):
            a.src, a.dst, a.absmax = src.t.data_ptr(), sp.buf.data_ptr(), sp.absmax.data_ptr()
            a.rows, a.cols, a.rows_pad, a.cols_pad, a.slab_k = sp.rows, sp.cols, sp.rows, sp.cols_pad, sp.slab_k
            a.absmax_given = 1 if given else 2          # 2: to be computed, slot already zero (no memset node)
        nb = sum((8.0 if given else 12.0) * sp.rows * sp.cols for _, sp, given in part)
        if _SPLIT_LOG is not None:
            _SPLIT_LOG.extend((sp.rows, sp.cols, given) for _, sp, given in part)
        _lib.call("eqv2_split_f16", ctypes.cast(arr, ctypes.c_void_p), len(part), _lib.stream_ptr(),
                  n_kernels=1 if all(g for _, _, g in part) else 2, work=(0.0, nb))
    if _SPLIT_SCOPES:
        for src, sp, _ in missing:
            _SPLIT_SCOPES[-1][src.key] = (src.t, sp)
    return out


def _split_of(splits, src):
    return splits.get(src.key) if src is not None else None


F16_MIN_MACS = 1 << 21      # launches below this many multiply-adds stay on the FFMA engine (tests set it to 0)


def _f16_forced(descs):
    return any(src is not None and isinstance(src.t, PlaneRef) for d in descs for src in d.src)


def _f16_ok(descs):
    """Should the f16x3 engine take this launch?  Measured (profiles/): the persistent TMA kernel beats the FFMA engine
    and the in-kernel-split tf32 engine from ~2 M multiply-adds per launch, operand splits included (a 640 x 128 x 128
    node-level linear takes 69 us on the FFMA engine: 5 CTAs).  Operands that exist only as planes force it."""
    if not all(_f16_addressable(d) for d in descs):
        assert not _f16_forced(descs), "an operand that exists only as fp16 planes is not addressable by the f16 engine"
        return False
    return _f16_forced(descs) or sum(d.M * d.N * d.K for d in descs) >= F16_MIN_MACS


def _f16_addressable(d):
    """Operands that map onto split buffers (contiguous fp32 tensors; block origin keeping TMA's 16-byte alignment)."""
    if min(d.a_rpb, d.b_rpb) < (1 << 31) and (d.src[0] is None or d.src[1] is None):
        return False
    for src in d.src:
        if src is None or not src.t.is_contiguous() or src.t.dtype != _F32 or src.col_off % 8 or src.col_off >= max(src.cols, 1):
            return False
    return min(d.M, d.N, d.K) > 0


_GEMM16_LOG = None        # diagnostics: set to a list to record (shapes, split_k) of every f16x3 launch, in order


def _run_gemm_f16(descs, split_k, flops, nbytes, out=None):
    if _GEMM16_LOG is not None:
        _GEMM16_LOG.append((tuple((d.M, d.N, d.K, d.transA, d.transB) for d in descs), split_k))
    splits = _splits_for([src for d in descs for src in d.src])
    n = len(descs)
    arr = (_lib.Gemm16Desc * n)()
    for a, d in zip(arr, descs):
        sa, sb = splits[d.src[0].key], splits[d.src[1].key]
        a.A = sa.buf.data_ptr() + 2 * (d.src[0].row_off * sa.cols_pad + d.src[0].col_off)
        a.B = sb.buf.data_ptr() + 2 * (d.src[1].row_off * sb.cols_pad + d.src[1].col_off)
        a.C, a.bias = d.C, d.bias
        a.a_absmax, a.b_absmax = sa.absmax.data_ptr(), sb.absmax.data_ptr()
        a.a_ld, a.a_plane, a.b_ld, a.b_plane, a.c_ld = sa.cols_pad, sa.plane, sb.cols_pad, sb.plane, d.c_ld
        a.c_rpb, a.c_bs = d.c_rpb, d.c_bs
        a.M, a.N, a.K, a.transA, a.transB, a.accumulate = d.M, d.N, d.K, d.transA, d.transB, d.accumulate
    slot = None
    if out is not None and split_k == 1 and _FEATURES["c_absmax"]:
        # the epilogue reduces max |C| while storing: whichever kernel consumes `out` next (a producer that writes operand
        # planes, or the operand split) finds it in the registry instead of re-reading the tensor
        slot = _zeroed_slot(out.device)
        for a in arr:
            a.c_absmax = slot.data_ptr()
    if _GEMM_MODE["mode"] == "f16":
        _lib.call("eqv2_gemm_f16_ex", ctypes.cast(arr, ctypes.c_void_p), n, int(split_k), 1, _lib.stream_ptr(),
                  work=(flops, nbytes))
    else:
        _lib.call("eqv2_gemm_f16", ctypes.cast(arr, ctypes.c_void_p), n, int(split_k), _lib.stream_ptr(),
                  work=(flops, nbytes))
    if slot is not None:
        _register_absmax(out, slot)
    return splits


_GEMM_LOG = None          # diagnostics: set to a list to record the launches the f16x3 engine declines


def run_gemm(descs, split_k=1, out=None):
    """-> {OperandSrc.key: SplitF16} of the operand splits the f16x3 engine used ({} on the other engines).
    `out`: the tensor all groups write into (when every group's C lies in one tensor): its max |.| is then reduced by the
    GEMM epilogue and registered (`_known_absmax`)."""
    n = len(descs)
    assert 1 <= n <= _lib.MAX_GEMM_GROUPS
    flops = sum(2.0 * d.M * d.N * d.K for d in descs)
    nbytes = sum(4.0 * (d.M * d.K + d.K * d.N + d.M * d.N) for d in descs)
    mode = _GEMM_MODE["mode"]
    if mode in ("f16x3", "f16"):
        if _f16_ok(descs):
            return _run_gemm_f16(descs, split_k, flops, nbytes, out)
        if _GEMM_LOG is not None:
            _GEMM_LOG.append([(d.M, d.N, d.K, d.transA, d.transB, _f16_addressable(d)) for d in descs])
        mode = "tf32x3"
    arr = (_lib.GemmDesc * n)(*descs)
    if mode != "fp32" and all(_tc_ok(d) for d in descs):
        _lib.call("eqv2_gemm_tc", ctypes.cast(arr, ctypes.c_void_p), n, int(split_k), 0 if mode == "tf32x3" else 1,
                  _lib.stream_ptr(), work=(flops, nbytes))
    else:
        _lib.call("eqv2_gemm_f32", ctypes.cast(arr, ctypes.c_void_p), n, int(split_k), _lib.stream_ptr(),
                  work=(flops, nbytes))
    return {}


# The dense contractions are four primitives that are CLOSED UNDER DIFFERENTIATION: the backward of each is
# a combination of the same primitives applied through autograd, so a backward pass run with
# create_graph=True (forces = -dE/dpos, train_MatPES_GATAWandB.py:72-77) can itself be differentiated
# (loss.backward() through the force computation) without any eager fallback.
#   SliceMm    : Y[:, ys_g] = X[:, xs_g] @ (W_g^T | W_g)      column-sliced operands (SO(2) blocks, nn.Linear)
#   SliceOuter : W_g = U[:, us_g]^T @ V[:, vs_g]               (weight gradients; split-K over the edge dimension)
#   SlabMm     : y[:, slab_l, :] = x[:, slab_l, :] @ (W_l^T | W_l)   degree slabs of [N, K, C] (SO3_LinearV2)
#   SlabOuter  : W_l = U[:, slab_l, :]^T @ V[:, slab_l, :]
def _covers(slices, width):
    pos = 0
    for off, w in sorted(slices):
        if off != pos:
            return False
        pos = off + w
    return pos == width


def _slice_mm(X, bias, xs, ys, y_width, transW, Ws):
    """Y[:, ys_g] = X[:, xs_g] @ (W_g^T | W_g) as ONE grouped launch -> (Y, [split of X, splits of the W_g]).
    X may be a PlaneRef (operand that exists only as fp16 planes, found through the active split_scope)."""
    rows, xw = X.shape
    Y = (torch.empty if _covers(ys, y_width) else torch.zeros)(rows, y_width, dtype=_F32, device=X.device)
    descs = []
    for g, ((xo, k), (yo, n)) in enumerate(zip(xs, ys)):
        assert tuple(Ws[g].shape) == ((n, k) if transW else (k, n)), (Ws[g].shape, n, k, transW)
        descs.append(_desc(X, Ws[g], Y, bias if g == 0 else None, rows, n, k, 0, 1 if transW else 0,
                           _plain(xw), _plain(Ws[g].shape[1]), _plain(y_width), a_off=xo, c_off=yo))
    sp = run_gemm(descs, out=Y) if rows > 0 else {}
    return Y, [_split_of(sp, descs[0].src[0])] + [_split_of(sp, d.src[1]) for d in descs]


def _slice_outer(U, V, us, vs):
    """W_g = U[:, us_g]^T @ V[:, vs_g] (split-K over the rows) -> (list of W_g, [split of U, split of V]); U / V may be
    PlaneRefs."""
    rows = U.shape[0]
    descs = []
    for (uo, m), (vo, n) in zip(us, vs):
        descs.append(_desc(U, V, U, None, m, n, rows, 1, 0, _plain(U.shape[1]), _plain(V.shape[1]), _plain(n),
                           a_off=uo, b_off=vo))
    split = _pick_split(descs, True) if rows > 0 else 1
    # all groups' outputs come out of ONE allocation (one zero-fill launch when split-K atomics need it, not one per
    # group); every block starts 16-byte aligned
    sizes = [(m * n + 3) // 4 * 4 for (_, m), (_, n) in zip(us, vs)]
    flat = (torch.zeros if (split > 1 or rows == 0) else torch.empty)(sum(sizes), dtype=_F32, device=U.device)
    outs, off = [], 0
    for d, sz, ((uo, m), (vo, n)) in zip(descs, sizes, zip(us, vs)):
        W = flat[off:off + m * n].view(m, n)
        off += sz
        d.C = W.data_ptr()
        outs.append(W)
    sp = run_gemm(descs, split) if rows > 0 else {}
    return outs, ([_split_of(sp, descs[0].src[0]), _split_of(sp, descs[0].src[1])] if descs else [None, None])


def _grad_wanted(t):
    """Will the autograd engine consume a gradient for `t` in the backward pass that is running NOW?  `ctx.needs_input_grad`
    only says that `t` requires grad at all; in the force pass of the MatPES family (`autograd.grad(E, pos,
    create_graph=True)`, train_MatPES_GATAWandB.py:72-77) no parameter gradient is wanted, and computing the weight-gradient
    GEMMs there anyway (a third of that pass) would be thrown away -- the reference's pure-autograd path prunes them too.
    The engine knows: it executes a node only if some requested input lies behind it.  Conservative on any doubt."""
    if t is None or not t.requires_grad:
        return False
    try:
        return bool(torch._C._will_engine_execute_node(torch.autograd.graph.get_gradient_edge(t).node))
    except Exception:       # leaf passed as an explicit autograd.grad input, API missing, not inside a backward pass, ...
        return True


class SliceMm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, bias, xs, ys, y_width, transW, *Ws):
        _lib.check_device(X, bias, *Ws)
        assert X.is_contiguous() and all(w.is_contiguous() for w in Ws)
        Y, ctx.splits = _slice_mm(X, bias, xs, ys, y_width, transW, Ws)
        ctx.save_for_backward(X, *Ws)
        ctx.spec = (xs, ys, y_width, transW, bias is not None)
        ctx.bias_wanted = (lambda b=bias: _grad_wanted(b))       # the bias is only needed for this query
        return Y

    @staticmethod
    def backward(ctx, gY):
        X, *Ws = ctx.saved_tensors
        xs, ys, y_width, transW, has_bias = ctx.spec
        gX = gb = None
        gWs = [None] * len(Ws)
        gY = gY.contiguous()
        with split_scope(zip([X, *Ws], ctx.splits)):      # forward splits of X / W; gY's is made once and shared
            if ctx.needs_input_grad[0] and _grad_wanted(X):
                gX = SliceMm.apply(gY, None, ys, xs, X.shape[1], not transW, *Ws)
            if any(ctx.needs_input_grad[6:]) and any(_grad_wanted(W) for W in Ws):
                if transW:      # W_g [n,k]: gW = gY_g^T X_g
                    outs = SliceOuter.apply(gY, X, ys, xs)
                else:           # W_g [k,n]: gW = X_g^T gY_g
                    outs = SliceOuter.apply(X, gY, xs, ys)
                gWs = list(outs) if isinstance(outs, tuple) else [outs]
        if has_bias and ctx.needs_input_grad[1] and ctx.bias_wanted():
            yo, n = ys[0]
            gb = colsum(gY, yo, n)
        return (gX, gb, None, None, None, None, *gWs)


class SliceOuter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, V, us, vs):
        _lib.check_device(U, V)
        assert U.is_contiguous() and V.is_contiguous()
        outs, ctx.splits = _slice_outer(U, V, us, vs)
        ctx.save_for_backward(U, V)
        ctx.spec = (us, vs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gWs):
        U, V = ctx.saved_tensors
        us, vs = ctx.spec
        gWs = [g.contiguous() if g is not None else torch.zeros(m, n, dtype=_F32, device=U.device)
               for g, ((_, m), (_, n)) in zip(gWs, zip(us, vs))]
        gU = gV = None
        with split_scope(zip([U, V], ctx.splits)):
            if ctx.needs_input_grad[0] and _grad_wanted(U):     # gU_g = V_g @ gW_g^T
                gU = SliceMm.apply(V, None, vs, us, U.shape[1], True, *gWs)
            if ctx.needs_input_grad[1] and _grad_wanted(V):     # gV_g = U_g @ gW_g
                gV = SliceMm.apply(U, None, us, vs, V.shape[1], False, *gWs)
        return gU, gV, None, None


class SlabMm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, bias, transW):
        _lib.check_device(x, W, bias)
        assert x.is_contiguous() and W.is_contiguous()
        N, K, Ci = x.shape
        L1, Wo, Wi = W.shape
        Co = Wo if transW else Wi
        assert Ci == (Wi if transW else Wo)
        y = torch.empty(N, K, Co, dtype=_F32, device=x.device)
        descs = []
        for l in range(L1):
            r = 2 * l + 1
            descs.append(_desc(x, W, y, bias if l == 0 else None, N * r, Co, Ci, 0, 1 if transW else 0,
                               (r, K * Ci, Ci), _plain(Wi), (r, K * Co, Co),
                               a_off=l * l * Ci, b_off=l * Wo * Wi, c_off=l * l * Co,
                               a_src=OperandSrc(x, N * K, Ci, K, N * l * l, 0),
                               b_src=OperandSrc(W, L1 * Wo, Wi, 0, l * Wo, 0)))
        sp = run_gemm(descs, out=y) if N > 0 else {}
        ctx.save_for_backward(x, W)
        ctx.spec = (transW, bias is not None)
        ctx.bias_wanted = (lambda b=bias: _grad_wanted(b))
        ctx.splits = [_split_of(sp, descs[0].src[0]), _split_of(sp, descs[0].src[1])]
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        transW, has_bias = ctx.spec
        gx = gW = gb = None
        gy = gy.contiguous()
        with split_scope(zip([x, W], ctx.splits)):
            if ctx.needs_input_grad[0] and _grad_wanted(x):
                gx = SlabMm.apply(gy, W, None, not transW)
            if ctx.needs_input_grad[1] and _grad_wanted(W):
                gW = SlabOuter.apply(gy, x) if transW else SlabOuter.apply(x, gy)
        if has_bias and ctx.needs_input_grad[2] and ctx.bias_wanted():
            # column sums of the l = 0 slab = the first C columns of gy read as [N, K*C] (the strided torch reduction took
            # 16 us per launch, 37 launches per OC20 step)
            gb = colsum(gy.view(gy.shape[0], -1), 0, gy.shape[2])
        return gx, gW, gb, None


class SlabOuter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, V):
        _lib.check_device(U, V)
        assert U.is_contiguous() and V.is_contiguous()
        N, K, Cu = U.shape
        Cv = V.shape[2]
        L1 = int(round(math.sqrt(K)))
        descs = []
        for l in range(L1):
            r = 2 * l + 1
            descs.append(_desc(U, V, U, None, Cu, Cv, N * r, 1, 0, (r, K * Cu, Cu), (r, K * Cv, Cv), _plain(Cv),
                               a_off=l * l * Cu, b_off=l * l * Cv,
                               a_src=OperandSrc(U, N * K, Cu, K, N * l * l, 0),
                               b_src=OperandSrc(V, N * K, Cv, K, N * l * l, 0)))
        split = _pick_split(descs, True) if N > 0 else 1
        W = (torch.zeros if (split > 1 or N == 0) else torch.empty)(L1, Cu, Cv, dtype=_F32, device=U.device)
        for l in range(L1):
            descs[l].C = W.data_ptr() + 4 * l * Cu * Cv
        sp = run_gemm(descs, split) if N > 0 else {}
        ctx.save_for_backward(U, V)
        ctx.splits = [_split_of(sp, descs[0].src[0]), _split_of(sp, descs[0].src[1])]
        return W

    @staticmethod
    def backward(ctx, gW):
        U, V = ctx.saved_tensors
        gU = gV = None
        gW = gW.contiguous()
        with split_scope(zip([U, V], ctx.splits)):
            if ctx.needs_input_grad[0] and _grad_wanted(U):     # gU_l = V_l @ gW_l^T
                gU = SlabMm.apply(V, gW, None, True)
            if ctx.needs_input_grad[1] and _grad_wanted(V):     # gV_l = U_l @ gW_l
                gV = SlabMm.apply(U, gW, None, False)
        return gU, gV


# Public entry points.  `.contiguous()` is applied HERE, as a differentiable torch op outside the Functions, so
# that the tensors a Function saves are the graph-connected ones (needed by the differentiable backward passes).
def linear(x, W, b=None):
    """y = x @ W^T + b  (nn.Linear inside radial_function.py:29, transformer_block.py:420)."""
    n, k = W.shape
    return SliceMm.apply(x.contiguous(), b, ((0, k),), ((0, n),), n, True, W.contiguous())


def so2_conv(A, bias0, groups, weights):
    """All m-blocks of one SO(2) convolution as ONE grouped GEMM (so2_ops.py:150-185, :53-61).
    A: [E, Kr*c_in] (m-primary, radial modulation already applied), weights[g]: [n_g, k_g] (m = 0: fc_m0.weight;
    m > 0: the 2x2 real block form [[Wr, -Wi], [Wi, Wr]] of the complex multiply), bias0: fc_m0.bias.
    groups[g] = (a_off, k_g, c_off, n_g).  Returns Y [E, extra + Kr*c_out]."""
    xs = tuple((a_off, k_g) for a_off, k_g, _, _ in groups)
    ys = tuple((c_off, n_g) for _, _, c_off, n_g in groups)
    width = sum(n for _, n in ys)
    return SliceMm.apply(A.contiguous(), bias0, xs, ys, width, True, *[w.contiguous() for w in weights])


def so3_linear(x, W, b):
    """SO3_LinearV2 (so3.py:698-743): one GEMM per degree l over the slab x[:, l^2:(l+1)^2, :], bias on l = 0."""
    return SlabMm.apply(x.contiguous(), W.contiguous(), b, True)


# ----------------------------------------------------------------------------------------------
# Wigner-D
# ----------------------------------------------------------------------------------------------
def wigner_from_rot(rot, lmax):
    """so3.py:525-545 -- block-diagonal packed [E, sum_l (2l+1)^2], no gradient (detached upstream)."""
    _lib.check_device(rot)
    rot = rot.detach().to(_F32).contiguous()
    E = rot.shape[0]
    lay = CoeffLayout.get(lmax, lmax)
    wig = torch.empty(E, lay.WS, dtype=_F32, device=rot.device)
    _lib.call("eqv2_wigner_from_rot", rot.data_ptr(), lay.dev(rot.device)["jd"].data_ptr(), wig.data_ptr(),
              E, lmax, _lib.stream_ptr())
    return wig


def edge_frames(edge_vec, draw=None):
    """Per-edge frames [E,3,3] (rows z, x_edge, -y), detached.  `draw` = the reference's helper draw
    (`torch.rand_like(vec) - 0.5`, edge_rot_mat.py:28) -> edge_rot_mat.py:13-80, with its two host-side conditions
    (lines 19-24: short-edge warning, line 58: assert) checked from one read-back; `draw=None` -> the deterministic frames
    of equiformerv2_MatPESv2.py:41-66 (no read-back: capturable in a CUDA graph)."""
    _lib.check_device(edge_vec, draw)
    vec = edge_vec.detach().to(_F32).contiguous()
    E = vec.shape[0]
    out = torch.empty(E, 3, 3, dtype=_F32, device=vec.device)
    if draw is None:
        _lib.call("eqv2_edge_frames", vec.data_ptr(), None, out.data_ptr(), E, 1, None, _lib.stream_ptr())
        return out
    draw = draw.detach().to(_F32).contiguous()
    stats = torch.zeros(2, dtype=torch.int32, device=vec.device)
    _lib.call("eqv2_edge_frames", vec.data_ptr(), draw.data_ptr(), out.data_ptr(), E, 0, stats.data_ptr(), _lib.stream_ptr())
    if E > 0:
        s0, s1 = stats.tolist()
        min_len = np.array([~s0 & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0]
        max_dot = np.array([s1 & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0]
        if min_len < 0.0001:
            print("Error edge_vec_0_distance: {}".format(min_len))
        assert max_dot < 0.99
    return out


def edge_sh(edge_vec, lmax):
    """Edge spherical harmonics l = 1..lmax in the original frame, detached (equiformerv2_MatPES_GATAV2.py:232-241)."""
    _lib.check_device(edge_vec)
    vec = edge_vec.detach().to(_F32).contiguous()
    E = vec.shape[0]
    out = torch.empty(E, (lmax + 1) ** 2 - 1, dtype=_F32, device=vec.device)
    _lib.call("eqv2_edge_sh", vec.data_ptr(), out.data_ptr(), E, lmax, _lib.stream_ptr())
    return out


def wigner_to_dense(wig, lmax):
    """Expand the packed blocks to the reference's dense [E,K,K] layout (diagnostics / API parity)."""
    E = wig.shape[0]
    K = (lmax + 1) ** 2
    out = wig.new_zeros(E, K, K)
    off = 0
    for l in range(lmax + 1):
        n = 2 * l + 1
        out[:, l * l:l * l + n, l * l:l * l + n] = wig[:, off:off + n * n].view(E, n, n)
        off += n * n
    return out


def _rot_work(lay, E, N, cols, node_cols, floats_per_edge):
    """Algorithmic (FLOPs, bytes) of one Wigner rotate / inverse-rotate launch: per edge and feature column every kept
    row of degree l costs 2l+1 multiply-adds (block-diagonal rotation) + 1 for the radial / attention weighting; traffic =
    the per-edge operands read / written once (`floats_per_edge`) + one pass over the node tensor [N, K, node_cols]."""
    macs = sum(2 * l + 2 for l in lay.row_l)
    return (2.0 * E * cols * macs, 4.0 * (E * floats_per_edge + N * lay.K * node_cols))


def _gr_fwd(x, rad, plan, wig, lmax, mmax):
    lay = CoeffLayout.get(lmax, mmax)
    tabs = lay.dev(x.device)
    N, K, C = x.shape
    nrad = lay.nslot * 2 * C
    if rad is not None:
        assert rad.shape == (plan.E, nrad), (rad.shape, plan.E, nrad)
    out = torch.empty(plan.E, lay.Kr * 2 * C, dtype=_F32, device=x.device)
    slot = _absmax_slot(x.device)
    _lib.call("eqv2_gather_rotate_fwd", x.data_ptr(), plan.src.data_ptr(), plan.dst.data_ptr(), wig.data_ptr(),
              _lib.ptr(rad), out.data_ptr(), tabs["pos_of_full"].data_ptr(), tabs["rad_slot"].data_ptr(),
              plan.E, C, lmax, mmax, lay.Kr, nrad, _lib.ptr(slot), _lib.stream_ptr(),
              work=_rot_work(lay, plan.E, N, 2 * C, C, (nrad if rad is not None else 0) + wig.shape[1] + lay.Kr * 2 * C))
    _register_absmax(out, slot)
    return out


def _gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_dx=True, want_drad=True):
    """(dx, drad) of T(x, rad, g) = <g, rad * (W x)>:  dx = sum_e W^t (g*rad) uses (rad, g);  drad = g*(W x) uses (x, g)."""
    lay = CoeffLayout.get(lmax, mmax)
    N, K, C = x.shape
    nrad = lay.nslot * 2 * C
    gx = grad = None
    if want_dx:
        gx = torch.empty_like(x)
        _lib.call("eqv2_gather_rotate_dx", wig.data_ptr(), _lib.ptr(rad), gA.data_ptr(), plan.rowptr_src.data_ptr(),
                  plan.perm_src.data_ptr(), plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), gx.data_ptr(), N, C,
                  lmax, mmax, lay.Kr, nrad, _lib.stream_ptr(),
                  work=_rot_work(lay, plan.E, N, 2 * C, C, (nrad if rad is not None else 0) + wig.shape[1] + lay.Kr * 2 * C))
    if want_drad:
        grad = torch.empty(plan.E, nrad, dtype=_F32, device=x.device)
        slot = _absmax_slot(x.device)
        _lib.call("eqv2_gather_rotate_drad", x.data_ptr(), plan.src.data_ptr(), plan.dst.data_ptr(), wig.data_ptr(),
                  gA.data_ptr(), grad.data_ptr(), plan.E, C, lmax, mmax, lay.Kr, nrad, _lib.ptr(slot), _lib.stream_ptr(),
                  work=_rot_work(lay, plan.E, N, 2 * C, C, nrad + wig.shape[1] + lay.Kr * 2 * C))
        _register_absmax(grad, slot)
    return gx, grad


class GatherRotateFn(torch.autograd.Function):
    """x[src] | x[dst] -> Wigner rotate -> |m|<=mmax rows, m-primary -> * radial weights.
    (transformer_block.py:250-275, so3.py:343-360, so3.py:322-334, so2_ops.py:150-175)
    Trilinear form T(x, rad, g); forward = dT/dg, backward = (dT/dx, dT/drad), and the derivative of the
    backward is again made of the same two kernels (GatherRotateBwdFn)."""

    @staticmethod
    def forward(ctx, x, rad, plan, wig, lmax, mmax):
        _lib.check_device(x, rad, wig)
        assert x.is_contiguous() and (rad is None or rad.is_contiguous())
        out = _gr_fwd(x, rad, plan, wig, lmax, mmax)
        ctx.save_for_backward(x, rad, wig)
        ctx.plan, ctx.lm = plan, (lmax, mmax)
        return out

    @staticmethod
    def backward(ctx, gA):
        x, rad, wig = ctx.saved_tensors
        gx, grad = GatherRotateBwdFn.apply(x, rad, gA.contiguous(), ctx.plan, wig, *ctx.lm)
        return gx, grad, None, None, None, None


class GatherRotateBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rad, gA, plan, wig, lmax, mmax):
        assert gA.is_contiguous()
        gx, grad = _gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_drad=rad is not None)
        ctx.save_for_backward(x, rad, gA, wig)
        ctx.plan, ctx.lm = plan, (lmax, mmax)
        if grad is None:
            ctx.mark_non_differentiable()
        return gx, grad

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, r):
        """cotangents u (of gx), r (of grad):  S = T(u, rad, g) + T(x, r, g)."""
        x, rad, gA, wig = ctx.saved_tensors
        plan, (lmax, mmax) = ctx.plan, ctx.lm
        d_x = d_rad = d_g = None
        if r is not None and rad is not None:
            r = r.contiguous()
            d_x, _ = _gr_bwd(x, r, gA, plan, wig, lmax, mmax, want_drad=False)
            d_g = _gr_fwd(x, r, plan, wig, lmax, mmax)
        if u is not None:
            u = u.contiguous()
            if rad is not None:
                _, d_rad = _gr_bwd(u, rad, gA, plan, wig, lmax, mmax, want_dx=False, want_drad=True)
            t = _gr_fwd(u, rad, plan, wig, lmax, mmax)
            d_g = t if d_g is None else d_g + t
        return d_x, d_rad, d_g, None, None, None, None


def _rir_fwd(val, alpha, plan, wig, lmax, mmax, rows_used, heads, scale, Cv):
    lay = CoeffLayout.get(lmax, mmax)
    tabs = lay.dev(val.device)
    out = torch.empty(plan.N, lay.K, Cv, dtype=_F32, device=val.device)
    _lib.call("eqv2_rotinv_reduce_fwd", val.data_ptr(), _lib.ptr(alpha), wig.data_ptr(),
              plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), out.data_ptr(),
              tabs["pos_of_full"].data_ptr(), plan.N, Cv, rows_used, rows_used * Cv, heads, lmax, mmax,
              float(scale), _lib.stream_ptr(),
              work=_rot_work(lay, plan.E, plan.N, Cv, Cv, rows_used * Cv + heads + wig.shape[1]))
    return out


def _rir_bwd(gout, val, alpha, plan, wig, lmax, mmax, rows_used, heads, scale, Cv):
    """(dval, dalpha) of T(val, alpha, g):  dval = alpha * (W g) uses (alpha, g);  dalpha = <W g, val> uses (val, g)."""
    lay = CoeffLayout.get(lmax, mmax)
    tabs = lay.dev(val.device)
    gval = torch.empty_like(val)
    galpha = torch.empty_like(alpha) if alpha is not None else None
    slot = _absmax_slot(val.device)
    _lib.call("eqv2_rotinv_reduce_bwd", gout.data_ptr(), val.data_ptr(), _lib.ptr(alpha), wig.data_ptr(),
              plan.dst.data_ptr(), gval.data_ptr(), _lib.ptr(galpha), tabs["pos_of_full"].data_ptr(),
              plan.E, Cv, rows_used, rows_used * Cv, heads, lmax, mmax, float(scale), _lib.ptr(slot), _lib.stream_ptr(),
              work=_rot_work(lay, plan.E, plan.N, Cv, Cv, 2 * rows_used * Cv + 2 * heads + wig.shape[1]))
    _register_absmax(gval, slot)
    return gval, galpha


class RotInvReduceFn(torch.autograd.Function):
    """(value * alpha) -> Wigner^T with l>mmax rescale -> deterministic dst-segmented sum.
    (transformer_block.py:321-331, so3.py:367-387,516-521, so3.py:304-318; input_block.py:113-129)
    Trilinear form T(val, alpha, g) like GatherRotateFn."""

    @staticmethod
    def forward(ctx, val, alpha, plan, wig, lmax, mmax, rows_used, heads, scale):
        _lib.check_device(val, alpha, wig)
        assert val.is_contiguous() and (alpha is None or alpha.is_contiguous())
        Cv = val.shape[1] // rows_used
        meta = (lmax, mmax, rows_used, heads, float(scale), Cv)
        out = _rir_fwd(val, alpha, plan, wig, *meta)
        ctx.save_for_backward(val, alpha, wig)
        ctx.plan, ctx.meta = plan, meta
        return out

    @staticmethod
    def backward(ctx, gout):
        val, alpha, wig = ctx.saved_tensors
        gval, galpha = RotInvReduceBwdFn.apply(gout.contiguous(), val, alpha, ctx.plan, wig, ctx.meta)
        return gval, galpha, None, None, None, None, None, None, None


class RotInvReduceBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gout, val, alpha, plan, wig, meta):
        assert gout.is_contiguous()
        gval, galpha = _rir_bwd(gout, val, alpha, plan, wig, *meta)
        ctx.save_for_backward(gout, val, alpha, wig)
        ctx.plan, ctx.meta = plan, meta
        return gval, galpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, a):
        """cotangents u (of gval), a (of galpha):  S = T(u, alpha, g) + T(val, a, g)."""
        gout, val, alpha, wig = ctx.saved_tensors
        plan, meta = ctx.plan, ctx.meta
        d_g = d_val = d_alpha = None
        if u is not None:
            u = u.contiguous()
            d_g = _rir_fwd(u, alpha, plan, wig, *meta)
            if alpha is not None:
                _, d_alpha = _rir_bwd(gout, u, alpha, plan, wig, *meta)
        if a is not None and alpha is not None:
            a = a.contiguous()
            t = _rir_fwd(val, a, plan, wig, *meta)
            d_g = t if d_g is None else d_g + t
            d_val, _ = _rir_bwd(gout, val, a, plan, wig, *meta)
        return d_g, d_val, d_alpha, None, None, None


# ----------------------------------------------------------------------------------------------
# First-order fused edge path of the f16 engine: producer kernels write the GEMM operand planes themselves
# ----------------------------------------------------------------------------------------------
# With the f16x3 engine every fp32 GEMM operand used to be re-read and re-written once by the operand split (r01:
# 12.7 % + 2.3 % of the OC20 step, 24.5 GB of HBM traffic per step).  The three edge-parallel rotate kernels now write
# scaled fp16 hi/lo planes directly, with the scale taken from a bound computed on the device from the registered
# maxima of their inputs (csrc/common.cuh).  Only for steps that are differentiated once (configs 1-2): the operand
# then never exists as an fp32 tensor autograd could differentiate again.
def _planes_for(rows, cols, device):
    """(PlaneRef, SplitF16) of a [rows, cols] operand a producer kernel is about to write.  cols % 8 == 0 (16-byte rows);
    the padding columns up to the next multiple of 64 stay unwritten -- the GEMM's tensor maps are bounded by the true
    extents, reads beyond them are zero-filled by the TMA unit."""
    ref = PlaneRef(rows, cols, device)
    return ref, SplitF16(OperandSrc(ref, rows, cols))


def fused_planes_available(x, h, W3, cols):
    """Can gather_rotate write conv1's A operand as planes?  f16 engine, whole warps of channels, the maximum of x known,
    and the radial output layer large enough to run on the f16 engine (its epilogue then supplies the maximum of rad)."""
    if not (_FEATURES["planes"] and _GEMM_MODE["mode"] in ("f16x3", "f16") and cols % 8 == 0 and x.is_cuda
            and (2 * x.shape[2]) % 32 == 0 and (2 * x.shape[2]) % min(256, 2 * x.shape[2]) == 0
            and hasattr(_lib.lib(), "eqv2_gather_rotate_fwd_planes") and _known_absmax(x) is not None):
        return False
    E, k3 = h.shape
    nrad = W3.shape[0]
    return (h.dtype == _F32 and W3.dtype == _F32 and k3 % 8 == 0 and nrad % 8 == 0 and _FEATURES["c_absmax"]
            and E * nrad * k3 >= F16_MIN_MACS and E > 0)


class GatherRotateConvFn(torch.autograd.Function):
    """last radial-MLP layer (radial_function.py:29) -> gather x[src]|x[dst] + Wigner rotate + radial modulation
    (transformer_block.py:250-275, so2_ops.py:142-175) -> first SO(2) convolution (so2_ops.py:150-185), with the two big
    per-edge intermediates living only as GEMM operand planes:
      * the rotated / modulated rows [E, Kr*2C]: written once by `eqv2_gather_rotate_fwd_planes`, read by the forward GEMM
        and again by the weight-gradient GEMM;
      * the gradient of the radial weights [E, n_rad]: written by `eqv2_gather_rotate_drad_planes`, read by the radial
        layer's dgrad / wgrad GEMMs and (for its bias gradient) by `eqv2_planes_colsum`.
    Backward: dgrad + wgrad GEMMs of the convolution, the dx kernel on the fp32 dgrad output, then the radial layer's two
    GEMMs.  Bounds come from the registered maxima of x (norm kernel), rad and dA (GEMM epilogues)."""

    @staticmethod
    def forward(ctx, x, h, W3, b3, bias0, plan, wig, lmax, mmax, groups, *Ws):
        _lib.check_device(x, h, W3, b3, wig, bias0, *Ws)
        assert x.is_contiguous() and h.is_contiguous() and W3.is_contiguous() and all(w.is_contiguous() for w in Ws)
        lay = CoeffLayout.get(lmax, mmax)
        N, K, C = x.shape
        E, cols, nrad = plan.E, lay.Kr * 2 * C, lay.nslot * 2 * C
        assert W3.shape == (nrad, h.shape[1]) and h.shape[0] == E, (W3.shape, h.shape, E, nrad)
        rad, rsplits = _slice_mm(h, b3, ((0, h.shape[1]),), ((0, nrad),), nrad, True, [W3])
        bx, brad = _known_absmax(x), _known_absmax(rad)
        assert bx is not None and brad is not None, "GatherRotateConvFn needs the registered maxima of x and rad"
        ref, spA = _planes_for(E, cols, x.device)
        _lib.call("eqv2_gather_rotate_fwd_planes", x.data_ptr(), plan.src.data_ptr(), plan.dst.data_ptr(), wig.data_ptr(),
                  rad.data_ptr(), spA.buf.data_ptr(), spA.plane, spA.cols_pad, bx.data_ptr(), brad.data_ptr(),
                  spA.absmax.data_ptr(), E, C, lmax, mmax, lay.Kr, nrad, _lib.stream_ptr(),
                  work=_rot_work(lay, E, N, 2 * C, C, nrad + wig.shape[1] + lay.Kr * C))      # planes: 2 x 2 B per value
        xs = tuple((a_off, k_g) for a_off, k_g, _, _ in groups)
        ys = tuple((c_off, n_g) for _, _, c_off, n_g in groups)
        width = sum(n for _, n in ys)
        with split_scope([(ref, spA)]):
            Y, splits = _slice_mm(ref, bias0, xs, ys, width, True, Ws)
        ctx.save_for_backward(x, h, W3, rad, wig, *Ws)
        ctx.plan, ctx.lm, ctx.spec = plan, (lmax, mmax), (xs, ys, width, bias0 is not None, b3 is not None)
        ctx.ref, ctx.spA, ctx.wsplits, ctx.rsplits = ref, spA, splits[1:], rsplits
        return Y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gY):
        x, h, W3, rad, wig, *Ws = ctx.saved_tensors
        xs, ys, width, has_bias, has_b3 = ctx.spec
        plan, (lmax, mmax) = ctx.plan, ctx.lm
        lay = CoeffLayout.get(lmax, mmax)
        N, K, C = x.shape
        E, nrad = plan.E, rad.shape[1]
        gY = gY.contiguous()
        need_w = any(ctx.needs_input_grad[11:])
        with split_scope([(ctx.ref, ctx.spA)] + list(zip(Ws, ctx.wsplits))):     # gY is split once for both products
            gA, _ = _slice_mm(gY, None, ys, xs, ctx.ref.shape[1], False, Ws)
            gWs = _slice_outer(gY, ctx.ref, ys, xs)[0] if need_w else [None] * len(Ws)
        gb = colsum(gY, ys[0][0], ys[0][1]) if (has_bias and ctx.needs_input_grad[4]) else None
        gx = None
        if ctx.needs_input_grad[0]:
            gx, _ = _gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_dx=True, want_drad=False)
        gh = gW3 = gb3 = None
        if any(ctx.needs_input_grad[1:4]):
            bx, bg = _known_absmax(x), _known_absmax(gA)
            k3 = h.shape[1]
            scope = list(zip([h, W3], ctx.rsplits))
            if (bx is not None and bg is not None and _FEATURES["planes"] and nrad % 8 == 0 and
                    hasattr(_lib.lib(), "eqv2_gather_rotate_drad_planes")):
                dref, spD = _planes_for(E, nrad, x.device)
                _lib.call("eqv2_gather_rotate_drad_planes", x.data_ptr(), plan.src.data_ptr(), plan.dst.data_ptr(),
                          wig.data_ptr(), gA.data_ptr(), spD.buf.data_ptr(), spD.plane, spD.cols_pad, bx.data_ptr(),
                          bg.data_ptr(), spD.absmax.data_ptr(), E, C, lmax, mmax, lay.Kr, nrad, _lib.stream_ptr(),
                          work=_rot_work(lay, E, N, 2 * C, C, nrad // 2 + wig.shape[1] + lay.Kr * 2 * C))
                d_op = dref
                scope.append((dref, spD))
                if has_b3 and ctx.needs_input_grad[3]:
                    S = max(1, min(E // 16, -(-2368 // ((nrad + 127) // 128))))
                    partial = torch.empty(S * nrad, dtype=_F32, device=x.device)
                    gb3 = torch.empty(nrad, dtype=_F32, device=x.device)
                    _lib.call("eqv2_planes_colsum", spD.buf.data_ptr(), spD.plane, spD.cols_pad, 0, E, nrad, S,
                              spD.absmax.data_ptr(), partial.data_ptr(), gb3.data_ptr(), _lib.stream_ptr(), n_kernels=2,
                              work=(0.0, 4.0 * E * nrad))
            else:
                _, d_op = _gr_bwd(x, rad, gA, plan, wig, lmax, mmax, want_dx=False, want_drad=True)
                if has_b3 and ctx.needs_input_grad[3]:
                    gb3 = colsum(d_op, 0, nrad)
            with split_scope(scope):
                if ctx.needs_input_grad[1]:
                    gh, _ = _slice_mm(d_op, None, ((0, nrad),), ((0, k3),), k3, False, [W3])
                if ctx.needs_input_grad[2]:
                    gW3 = _slice_outer(d_op, h, ((0, nrad),), ((0, k3),))[0][0]
        return (gx, gh, gW3, gb3, gb, None, None, None, None, None, *gWs)


class ConvRotInvReduceFn(torch.autograd.Function):
    """second SO(2) convolution (transformer_block.py:305) -> alpha-weighting + inverse rotation + deterministic
    dst-segmented sum (transformer_block.py:321-331, so3.py:367-387,304-318).  Backward: `eqv2_rotinv_reduce_bwd_planes`
    writes d(value) directly as the operand planes of the dgrad / wgrad GEMMs (bound from the registered maximum of the
    node gradient and `alpha_bound` >= max |alpha|); Z is kept only as the planes the forward GEMM read."""

    @staticmethod
    def forward(ctx, Zm, alpha, bias0, plan, wig, lmax, mmax, heads, alpha_bound, groups, *Ws):
        _lib.check_device(Zm, alpha, wig, bias0, *Ws)
        given = getattr(Zm, "_eqv2_planes", None)       # Z exists only as planes (edge_act_alpha(z_planes=True))
        assert (given is not None or Zm.is_contiguous()) and alpha.is_contiguous() and all(w.is_contiguous() for w in Ws)
        lay = CoeffLayout.get(lmax, mmax)
        xs = tuple((a_off, k_g) for a_off, k_g, _, _ in groups)
        ys = tuple((c_off, n_g) for _, _, c_off, n_g in groups)
        width = sum(n for _, n in ys)
        if given is not None:
            with split_scope([given]):
                V, splits = _slice_mm(given[0], bias0, xs, ys, width, True, Ws)
        else:
            V, splits = _slice_mm(Zm, bias0, xs, ys, width, True, Ws)
        Cv = width // lay.Kr
        meta = (lmax, mmax, lay.Kr, heads, 1.0, Cv)
        out = _rir_fwd(V, alpha, plan, wig, *meta)
        planes_ok = (_FEATURES["planes"] and splits[0] is not None and width % 8 == 0 and Cv % 32 == 0 and heads <= Cv
                     and hasattr(_lib.lib(), "eqv2_rotinv_reduce_bwd_planes"))
        assert planes_ok or given is None, "Z was written as planes only, but the backward pass would need it in fp32"
        if planes_ok:           # Z itself is not needed again: its planes serve the weight gradient
            ctx.save_for_backward(V, alpha, wig, *Ws)
            ctx.zref = given[0] if given is not None else PlaneRef(Zm.shape[0], Zm.shape[1], Zm.device)
            ctx.spZ = splits[0]
            ctx.spZ.version = 0
        else:
            ctx.save_for_backward(V, alpha, wig, *Ws, Zm)
            ctx.zref, ctx.spZ = None, splits[0]
        ctx.plan, ctx.meta, ctx.spec, ctx.wsplits = plan, meta, (xs, ys, width, float(alpha_bound)), splits[1:]
        ctx.nw, ctx.has_bias = len(Ws), bias0 is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        V, alpha, wig = saved[:3]
        Ws = saved[3:3 + ctx.nw]
        xs, ys, width, alpha_bound = ctx.spec
        plan, meta = ctx.plan, ctx.meta
        lmax, mmax, rows_used, heads, scale, Cv = meta
        lay = CoeffLayout.get(lmax, mmax)
        gout = gout.contiguous()
        bound = _known_absmax(gout)
        E = plan.E
        need_w = any(ctx.needs_input_grad[10:])
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        gb = None
        if ctx.zref is not None and bound is not None:
            gref, spG = _planes_for(E, width, gout.device)
            galpha = torch.empty_like(alpha)
            _lib.call("eqv2_rotinv_reduce_bwd_planes", gout.data_ptr(), V.data_ptr(), alpha.data_ptr(), wig.data_ptr(),
                      plan.dst.data_ptr(), spG.buf.data_ptr(), spG.plane, spG.cols_pad, bound.data_ptr(),
                      float(alpha_bound), spG.absmax.data_ptr(), galpha.data_ptr(), E, Cv, rows_used, rows_used * Cv,
                      heads, lmax, mmax, float(scale), _lib.stream_ptr(),
                      work=_rot_work(lay, E, plan.N, Cv, Cv, 2 * rows_used * Cv + 2 * heads + wig.shape[1]))
            scope = [(gref, spG), (ctx.zref, ctx.spZ)] + list(zip(Ws, ctx.wsplits))
            gv_op, z_op = gref, ctx.zref
            if want_b:          # bias gradient = column sums of d(value) over the m = 0 block, read from the planes
                yo, n = ys[0]
                S = max(1, min(E // 16, -(-2368 // ((n + 127) // 128))))
                partial = torch.empty(S * n, dtype=_F32, device=gout.device)
                gb = torch.empty(n, dtype=_F32, device=gout.device)
                _lib.call("eqv2_planes_colsum", spG.buf.data_ptr(), spG.plane, spG.cols_pad, yo, E, n, S,
                          spG.absmax.data_ptr(), partial.data_ptr(), gb.data_ptr(), _lib.stream_ptr(), n_kernels=2,
                          work=(0.0, 4.0 * E * n))
        else:
            # the node gradient's maximum is unknown (it did not come from an instrumented kernel) or Z was not split:
            # d(value) as an fp32 tensor (its maximum is reduced while it is written); Z from its planes if it has them
            gv_op, galpha = _rir_bwd(gout, V, alpha, plan, wig, *meta)
            z_op = ctx.zref if ctx.zref is not None else saved[3 + ctx.nw]
            scope = [(z_op, ctx.spZ)] + list(zip(Ws, ctx.wsplits))
            if want_b:
                gb = colsum(gv_op, ys[0][0], ys[0][1])
        with split_scope(scope):
            gZ, _ = _slice_mm(gv_op, None, ys, xs, z_op.shape[1], False, Ws)
            gWs = _slice_outer(gv_op, z_op, ys, xs)[0] if need_w else [None] * len(Ws)
        return (gZ, galpha, gb, None, None, None, None, None, None, None, *gWs)


def gather_rotate_conv(x, h, W3, b3, bias0, plan, wig, lmax, mmax, groups, weights):
    return GatherRotateConvFn.apply(x.contiguous(), h.contiguous(), W3.contiguous(), b3, bias0, plan, wig, lmax, mmax,
                                    groups, *[w.contiguous() for w in weights])


def conv_rotinv_reduce(Zm, alpha, bias0, plan, wig, lmax, mmax, heads, alpha_bound, groups, weights):
    if getattr(Zm, "_eqv2_planes", None) is None:
        Zm = Zm.contiguous()
    return ConvRotInvReduceFn.apply(Zm, alpha.contiguous(), bias0, plan, wig, lmax, mmax, heads, alpha_bound,
                                    groups, *[w.contiguous() for w in weights])


# ----------------------------------------------------------------------------------------------
# S2 activation + attention weights
# ----------------------------------------------------------------------------------------------
class GridMats:
    """Kernel-side form of one SO3_Grid: padded dense [G, KP] to/from-grid matrices in a given coefficient
    order (so3.py:584-622) and, when the grid is the resolution-18 one every config uses, the
    latitude/longitude factor tables of the separable kernel (csrc/s2act_sep.cu)."""

    def __init__(self, T, Fm, Kr, KP, G, lmax, mmax, order, factors):
        self.T, self.F, self.Kr, self.KP, self.G = T, Fm, Kr, KP, G
        self.lmax, self.mmax, self.order, self.factors = lmax, mmax, order, factors

    @classmethod
    def from_buffers(cls, to_grid, from_grid, lmax, mmax, order):
        """to_grid / from_grid: [res_beta, res_alpha, Kr] buffers of SO3_Grid (l-primary reduced)."""
        Kr = to_grid.shape[-1]
        G = to_grid.shape[0] * to_grid.shape[1]
        tg = to_grid.reshape(G, Kr).to(_F32)
        fg = from_grid.reshape(G, Kr).to(_F32)
        if order == "m":
            perm = torch.tensor(CoeffLayout.get(lmax, mmax).to_m, dtype=torch.long, device=tg.device)
            tg, fg = tg[:, perm], fg[:, perm]
        KP = _lib.lib().eqv2_s2act_padded_rows(Kr)
        if KP < 0:
            raise _lib.Eqv2Error(f"S2 activation: {Kr} coefficients unsupported")
        T = torch.zeros(G, KP, dtype=_F32, device=tg.device)
        Fm = torch.zeros(G, KP, dtype=_F32, device=tg.device)
        T[:, :Kr], Fm[:, :Kr] = tg, fg
        factors = cls._factor_tables(to_grid, from_grid, lmax, mmax)
        mats = cls(T.contiguous(), Fm.contiguous(), Kr, KP, G, lmax, mmax, order, factors)
        # operator-norm bound of the activation, for kernels that write its output as operand planes:
        # |F silu(T x)|_inf <= ||F^t||_inf ||T||_inf |x|_inf  (|silu(g)| <= |g|);  >= 1 covers the gate row silu(gate)
        mats.out_bound = 1.01 * max(1.0, float(tg.abs().sum(1).max()) * float(fg.abs().sum(0).max()))
        return mats

    @staticmethod
    def _factor_tables(to_grid, from_grid, lmax, mmax):
        """Flat fp32 block [Pt | Pf | cos | sin] if the buffers are the separable resolution-18 matrices
        (verified to 2e-6 against the buffers themselves), else None -> dense kernel."""
        res = to_grid.shape[0]
        if res != 18 or to_grid.shape[1] != 18 or lmax > 6 or not _lib.lib().eqv2_s2sep_supported(lmax, mmax):
            return None
        Pt, Pf, ct, st = _so3_math.s2_grid_factors(lmax, mmax, res)
        red = [(l, m) for l in range(lmax + 1) for m in range(-min(l, mmax), min(l, mmax) + 1)]
        tg, fg = to_grid.detach().cpu().numpy(), from_grid.detach().cpu().numpy()
        for i, (l, m) in enumerate(red):
            trig = ct[:, m] if m >= 0 else st[:, -m]
            if (np.abs(Pt[abs(m), :, l][:, None] * trig[None, :] - tg[:, :, i]).max() > 2e-6 or
                    np.abs(Pf[abs(m), :, l][:, None] * trig[None, :] - fg[:, :, i]).max() > 2e-6 * max(1.0, np.abs(fg).max())):
                return None
        return np.ascontiguousarray(np.concatenate([a.reshape(-1) for a in (Pt, Pf, ct, st)]).astype(np.float32))


_s2_slots = {}            # (device, lmax, mmax, order) -> [slot, host tables, last use, pinned]
_s2_clock = [0]
S2_SLOTS = 6


def _s2_bind_tables(mats, device):
    """Constant-memory slot holding `mats`' factor tables.  Each (device, lmax, mmax, coefficient order) in use owns one
    of S2_SLOTS slots; when they run out the least recently used one is rebound -- except slots that a CUDA-graph
    capture has used: those are PINNED for the life of the process, so a replayed graph can never see its tables
    replaced by another model's.  A miss during capture raises (the warm-up step binds everything)."""
    key = (str(device), mats.lmax, mats.mmax, mats.order)
    capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    _s2_clock[0] += 1
    hit = _s2_slots.get(key)
    if hit is not None:
        hit[2] = _s2_clock[0]
        hit[3] = hit[3] or capturing
        return hit[0]
    if capturing:
        raise _lib.Eqv2Error("S2 activation tables must be bound before a CUDA-graph capture (run one eager step first)")
    mine = {k: v for k, v in _s2_slots.items() if k[0] == key[0]}
    free = sorted(set(range(S2_SLOTS)) - {v[0] for v in mine.values()})
    if free:
        slot = free[0]
    else:
        victims = sorted((v[2], k) for k, v in mine.items() if not v[3])
        if not victims:
            raise _lib.Eqv2Error("S2 activation: all %d table slots are pinned by captured CUDA graphs" % S2_SLOTS)
        slot = _s2_slots.pop(victims[0][1])[0]
    _lib.call("eqv2_s2sep_set_tables", mats.factors.ctypes.data, int(mats.factors.size), slot, _lib.stream_ptr(),
              n_kernels=0)
    _s2_slots[key] = [slot, mats.factors, _s2_clock[0], False]       # the host block is kept alive
    return slot


def _s2_work(mats, R, C, passes):
    """Separable form (csrc/s2act_sep.cu): per (row, channel) and direction, the latitude transform costs 18 x Kr and the
    longitude transform 18 x 18 x (2 mmax + 1) multiply-adds; forward = to-grid + from-grid (`passes` = 2), backward
    recomputes the grid and applies both transposes (4), second order 6.  Bytes: the coefficient tensors in and out."""
    macs = 18 * mats.Kr + 324 * (2 * mats.mmax + 1)
    return (2.0 * R * C * macs * passes, 4.0 * R * C * (mats.Kr + 1) * (1 + passes // 2))


def _s2_fwd(mats, xp, x_rs, gp, g_rs, op, o_rs, R, C, device, absmax=None):
    if mats.factors is not None:
        slot = _s2_bind_tables(mats, device)
        _lib.call("eqv2_s2sep_fwd", xp, x_rs, gp, g_rs, op, o_rs, R, C, mats.lmax, mats.mmax, int(mats.order == "m"), slot,
                  _lib.ptr(absmax), _lib.stream_ptr(), work=_s2_work(mats, R, C, 2))
        return absmax
    else:
        _lib.call("eqv2_s2act_fwd", xp, x_rs, gp, g_rs, op, o_rs, mats.T.data_ptr(), mats.F.data_ptr(), R, C, mats.Kr,
                  mats.KP, mats.G, _s2_blocks(R, C), _lib.stream_ptr())


def _s2_bwd(mats, xp, x_rs, gp, g_rs, dop, o_rs, dxp, dx_rs, dgp, dg_rs, R, C, device, absmax=None):
    if mats.factors is not None:
        slot = _s2_bind_tables(mats, device)
        _lib.call("eqv2_s2sep_bwd", xp, x_rs, gp, g_rs, dop, o_rs, dxp, dx_rs, dgp, dg_rs, R, C, mats.lmax, mats.mmax,
                  int(mats.order == "m"), slot, _lib.ptr(absmax), _lib.stream_ptr(), work=_s2_work(mats, R, C, 4))
        return absmax
    else:
        _lib.call("eqv2_s2act_bwd", xp, x_rs, gp, g_rs, dop, o_rs, dxp, dx_rs, dgp, dg_rs, mats.T.data_ptr(),
                  mats.F.data_ptr(), R, C, mats.Kr, mats.KP, mats.G, _s2_blocks(R, C), _lib.stream_ptr())


def _s2_blocks(R, C):
    work = R * ((C + 63) // 64)
    return int(max(1, min(work, 148 * 2)))


def _s2_bwd2(mats, xp, x_rs, gp, g_rs, dop, o_rs, up, u_rs, wp, w_rs, d2xp, d2x_rs, d2gp, d2g_rs, d2op, d2o_rs, R, C, device):
    if mats.factors is None:
        raise _lib.Eqv2Error("S2 activation: second-order terms need the resolution-18 factorised grid")
    slot = _s2_bind_tables(mats, device)
    _lib.call("eqv2_s2sep_bwd2", xp, x_rs, gp, g_rs, dop, o_rs, up, u_rs, wp, w_rs, d2xp, d2x_rs, d2gp, d2g_rs, d2op,
              d2o_rs, R, C, mats.lmax, mats.mmax, int(mats.order == "m"), slot, _lib.stream_ptr(),
              work=_s2_work(mats, R, C, 6))


class S2ActFn(torch.autograd.Function):
    """SeparableS2Activation (transformer_block.py:442-447, activation.py:173-192):
    x [R,Kr,C], gate [R,C] (or None: plain S2Activation) -> [R,Kr,C]."""

    @staticmethod
    def forward(ctx, x, gate, mats):
        _lib.check_device(x, gate)
        assert x.is_contiguous() and (gate is None or gate.is_contiguous())
        R, Kr, C = x.shape
        assert Kr == mats.Kr
        out = torch.empty_like(x)
        _s2_fwd(mats, x.data_ptr(), Kr * C, _lib.ptr(gate), C, out.data_ptr(), Kr * C, R, C, x.device)
        ctx.save_for_backward(x, gate)
        ctx.mats = mats
        return out

    @staticmethod
    def backward(ctx, go):
        x, gate = ctx.saved_tensors
        gx, gg = S2ActBwdFn.apply(x, gate, go.contiguous(), ctx.mats)
        return gx, gg, None


class S2ActBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gate, go, mats):
        assert go.is_contiguous()
        R, Kr, C = x.shape
        gx = torch.empty_like(x)
        gg = torch.empty_like(gate) if gate is not None else None
        _s2_bwd(mats, x.data_ptr(), Kr * C, _lib.ptr(gate), C, go.data_ptr(), Kr * C, gx.data_ptr(), Kr * C,
                _lib.ptr(gg), C, R, C, x.device)
        ctx.save_for_backward(x, gate, go)
        ctx.mats = mats
        return gx, gg

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, w):
        x, gate, go = ctx.saved_tensors
        mats = ctx.mats
        R, Kr, C = x.shape
        u = u.contiguous() if u is not None else None
        w = w.contiguous() if (w is not None and gate is not None) else None
        d_x = torch.empty_like(x)
        d_go = torch.empty_like(go)
        d_gate = torch.empty_like(gate) if gate is not None else None
        _s2_bwd2(mats, x.data_ptr(), Kr * C, _lib.ptr(gate), C, go.data_ptr(), Kr * C, _lib.ptr(u), Kr * C, _lib.ptr(w), C,
                 d_x.data_ptr(), Kr * C, _lib.ptr(d_gate), C, d_go.data_ptr(), Kr * C, R, C, x.device)
        return d_x, d_gate, d_go, None


class EdgeActAlphaFn(torch.autograd.Function):
    """FIRST-ORDER fused path (no position gradients, configs 1-2).  Consumes the first SO(2) convolution's
    output Y[E, heads*ach + H + Kr*H] and produces
      * Z[E, Kr*H]   = SeparableS2Activation(gate = Y[:, heads*ach : heads*ach+H], Y[:, extra:])
      * alpha[E, heads] = segment_softmax_dst( alpha_dot . SmoothLeakyReLU(LayerNorm(Y[:, :heads*ach])) )
    (transformer_block.py:289-315).  One backward fills one dY buffer -- no zero-fill, no adds.
    When edge distances carry gradient (forces by autograd) the block uses S2ActFn + AttnAlphaFn instead,
    whose backward passes are differentiable."""

    @staticmethod
    def forward(ctx, Y, ln_w, ln_b, alpha_dot, plan, mats, heads, ach, H, z_planes=False):
        _lib.check_device(Y, ln_w, ln_b, alpha_dot)
        assert Y.is_contiguous()
        E, W = Y.shape
        extra = heads * ach + H
        Kr = mats.Kr
        assert W == extra + Kr * H
        yp = Y.data_ptr()
        if z_planes:
            # Z goes out as the operand planes of the second convolution (ConvRotInvReduceFn finds them through the
            # `_eqv2_planes` attribute edge_act_alpha attaches); the tensor autograd sees is a stride-0 stand-in of Z's shape
            by = _known_absmax(Y)
            assert by is not None, "edge_act_alpha(z_planes=True) needs the registered maximum of Y"
            ref, spZ = _planes_for(E, Kr * H, Y.device)
            slot = _s2_bind_tables(mats, Y.device)
            _lib.call("eqv2_s2sep_fwd_planes", yp + 4 * extra, W, yp + 4 * heads * ach, W, spZ.buf.data_ptr(), spZ.plane,
                      spZ.cols_pad, by.data_ptr(), float(mats.out_bound), spZ.absmax.data_ptr(), E, H, mats.lmax, mats.mmax,
                      int(mats.order == "m"), slot, _lib.stream_ptr(), work=_s2_work(mats, E, H, 2))
            Z = Y.new_empty(1).expand(E, Kr * H)
            EdgeActAlphaFn.last_planes = (ref, spZ)
        else:
            Z = torch.empty(E, Kr * H, dtype=_F32, device=Y.device)
            # max |Z| is reduced while Z is written: the operand split of the second convolution skips its absmax pass
            zslot = _s2_fwd(mats, yp + 4 * extra, W, yp + 4 * heads * ach, W, Z.data_ptr(), Kr * H, E, H, Y.device,
                            absmax=_absmax_slot(Y.device))
            _register_absmax(Z, zslot)
        logits = torch.empty(E, heads, dtype=_F32, device=Y.device)
        alpha = torch.empty(E, heads, dtype=_F32, device=Y.device)
        ln_w_c = ln_w.contiguous() if ln_w is not None else None
        ln_b_c = ln_b.contiguous() if ln_b is not None else None
        alpha_dot = alpha_dot.contiguous()
        _lib.call("eqv2_attn_alpha_fwd", yp, W, _lib.ptr(ln_w_c), _lib.ptr(ln_b_c), alpha_dot.data_ptr(),
                  plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), logits.data_ptr(), alpha.data_ptr(),
                  E, plan.N, heads, ach, 1e-5, _lib.stream_ptr(), n_kernels=2)
        ctx.save_for_backward(Y, ln_w_c, ln_b_c, alpha_dot, alpha)
        ctx.plan, ctx.mats, ctx.meta = plan, mats, (heads, ach, H)
        return Z, alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gZ, galpha):
        Y, ln_w, ln_b, alpha_dot, alpha = ctx.saved_tensors
        plan, mats = ctx.plan, ctx.mats
        heads, ach, H = ctx.meta
        E, W = Y.shape
        extra = heads * ach + H
        Kr = mats.Kr
        dev = Y.device
        gY = torch.empty_like(Y)
        yp, gp = Y.data_ptr(), gY.data_ptr()
        if gZ is None:
            gZ = torch.zeros(E, Kr * H, dtype=_F32, device=dev)
        gZ = gZ.contiguous()
        # the S2 backward and the attention backward fill disjoint column ranges of gY and reduce max |gY| into ONE slot
        gslot = _s2_bwd(mats, yp + 4 * extra, W, yp + 4 * heads * ach, W, gZ.data_ptr(), Kr * H, gp + 4 * extra, W,
                        gp + 4 * heads * ach, W, E, H, dev, absmax=_absmax_slot(dev))
        if galpha is None:
            galpha = torch.zeros_like(alpha)
        galpha = galpha.contiguous()
        g_lnw = _small_zeros(ln_w.shape, dev) if ln_w is not None else None
        g_lnb = _small_zeros(ln_b.shape, dev) if ln_b is not None else None
        g_dot = _small_zeros(alpha_dot.shape, dev)
        dlogits = torch.empty_like(alpha)
        _lib.call("eqv2_attn_alpha_bwd", yp, W, _lib.ptr(ln_w), _lib.ptr(ln_b), alpha_dot.data_ptr(),
                  plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), alpha.data_ptr(), galpha.data_ptr(),
                  dlogits.data_ptr(), gp, W, _lib.ptr(g_lnw), _lib.ptr(g_lnb), g_dot.data_ptr(), E, plan.N, heads,
                  ach, 1e-5, _lib.ptr(gslot), _lib.stream_ptr(), n_kernels=2)
        _register_absmax(gY, gslot)
        return gY, g_lnw, g_lnb, g_dot, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# small per-row operators (LayerNorm+SiLU of the radial MLP, attention logits + segment softmax, equivariant norms, RBF).
# Forward and first-order backward are kernels.  When the backward pass itself is being recorded (create_graph=True:
# forces by autograd), the backward runs as a differentiable operator of its own (`*BwdFn`) whose backward is a
# CLOSED-FORM second-order kernel (`eqv2_ln_silu_bwd2`, `eqv2_attn_alpha_bwd2`, `eqv2_equiv_norm_bwd2`, `eqv2_rbf_bwd2`;
# r02 -- round 1 re-expressed the operator with torch primitives here) for a cotangent of the INPUT gradient -- the case of
# the force loss, where the first backward only feeds the position gradient.  Only a cotangent of a PARAMETER gradient
# (differentiating weight gradients again: no reference script does) still takes the generic torch-expression route.
# ----------------------------------------------------------------------------------------------
def _recording(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _slr(x):
    return 0.6 * x + 0.4 * x * (2.0 * torch.sigmoid(x) - 1.0)


def _alpha_expr(heads, ach, eps, dst, N):
    def fn(Ya, ln_w, ln_b, alpha_dot):
        a = Ya.reshape(-1, heads, ach)
        if ln_w is not None:
            a = torch.nn.functional.layer_norm(a, (ach,), ln_w, ln_b, eps)
        logits = (_slr(a) * alpha_dot.unsqueeze(0)).sum(-1)
        idx = dst.view(-1, 1).expand_as(logits)
        mx = torch.full((N, heads), float("-inf"), dtype=logits.dtype, device=logits.device)
        mx = mx.scatter_reduce(0, idx, logits.detach(), reduce="amax", include_self=True)
        e = (logits - mx.gather(0, idx)).exp()
        ssum = torch.zeros(N, heads, dtype=logits.dtype, device=logits.device).scatter_add(0, idx, e)
        return e / (ssum.gather(0, idx) + 1e-16)
    return fn


class AttnAlphaFn(torch.autograd.Function):
    """alpha[E, heads] = segment_softmax_dst(alpha_dot . SmoothLeakyReLU(LayerNorm(Ya)))  (transformer_block.py:311-315)."""

    @staticmethod
    def forward(ctx, Ya, ln_w, ln_b, alpha_dot, plan, heads, ach):
        _lib.check_device(Ya, ln_w, ln_b, alpha_dot)
        assert Ya.is_contiguous() and alpha_dot.is_contiguous()
        E = Ya.shape[0]
        logits = torch.empty(E, heads, dtype=_F32, device=Ya.device)
        alpha = torch.empty(E, heads, dtype=_F32, device=Ya.device)
        _lib.call("eqv2_attn_alpha_fwd", Ya.data_ptr(), heads * ach, _lib.ptr(ln_w), _lib.ptr(ln_b), alpha_dot.data_ptr(),
                  plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), logits.data_ptr(), alpha.data_ptr(),
                  E, plan.N, heads, ach, 1e-5, _lib.stream_ptr(), n_kernels=2)
        ctx.save_for_backward(Ya, ln_w, ln_b, alpha_dot, alpha)
        ctx.plan, ctx.meta = plan, (heads, ach)
        return alpha

    @staticmethod
    def backward(ctx, galpha):
        Ya, ln_w, ln_b, alpha_dot, alpha = ctx.saved_tensors
        plan, (heads, ach) = ctx.plan, ctx.meta
        if _recording(Ya, ln_w, ln_b, alpha_dot, galpha):      # forces by autograd: differentiable backward
            g = AttnAlphaBwdFn.apply(Ya, ln_w, ln_b, alpha_dot, alpha, galpha.contiguous(), plan, heads, ach)
            if ln_w is None:
                return g[0], None, None, g[1], None, None, None
            return g[0], g[1], g[2], g[3], None, None, None
        gY, g_lnw, g_lnb, g_dot, _ = _attn_alpha_bwd(Ya, ln_w, ln_b, alpha_dot, alpha, galpha.contiguous(), plan, heads, ach)
        return gY, g_lnw, g_lnb, g_dot, None, None, None


def _attn_alpha_bwd(Ya, ln_w, ln_b, alpha_dot, alpha, galpha, plan, heads, ach):
    E = Ya.shape[0]
    gY = torch.empty_like(Ya)
    g_lnw = torch.zeros_like(ln_w) if ln_w is not None else None
    g_lnb = torch.zeros_like(ln_b) if ln_b is not None else None
    g_dot = torch.zeros_like(alpha_dot)
    dlogits = torch.empty_like(alpha)
    _lib.call("eqv2_attn_alpha_bwd", Ya.data_ptr(), heads * ach, _lib.ptr(ln_w), _lib.ptr(ln_b), alpha_dot.data_ptr(),
              plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), alpha.data_ptr(), galpha.data_ptr(),
              dlogits.data_ptr(), gY.data_ptr(), heads * ach, _lib.ptr(g_lnw), _lib.ptr(g_lnb), g_dot.data_ptr(),
              E, plan.N, heads, ach, 1e-5, None, _lib.stream_ptr(), n_kernels=2)
    return gY, g_lnw, g_lnb, g_dot, dlogits


class AttnAlphaBwdFn(torch.autograd.Function):
    """First-order backward of AttnAlphaFn as a differentiable operator of (Ya, ln_w, ln_b, alpha_dot, alpha, galpha).
    Its backward for a cotangent of gY is the closed-form kernel pair `eqv2_attn_alpha_bwd2`; `alpha` is an input here, so
    its cotangent flows on through AttnAlphaFn's ordinary backward.  Cotangents of the parameter gradients take the generic
    torch-expression route."""

    @staticmethod
    def forward(ctx, Ya, ln_w, ln_b, alpha_dot, alpha, galpha, plan, heads, ach):
        gY, g_lnw, g_lnb, g_dot, dlogits = _attn_alpha_bwd(Ya, ln_w, ln_b, alpha_dot, alpha, galpha, plan, heads, ach)
        ctx.save_for_backward(Ya, ln_w, ln_b, alpha_dot, alpha, galpha, dlogits)
        ctx.plan, ctx.meta = plan, (heads, ach)
        ctx.has_ln = ln_w is not None
        ctx.set_materialize_grads(False)
        if not ctx.has_ln:
            return gY, g_dot
        return gY, g_lnw, g_lnb, g_dot

    @staticmethod
    def backward(ctx, u, *rest):
        Ya, ln_w, ln_b, alpha_dot, alpha, galpha, dlogits = ctx.saved_tensors
        plan, (heads, ach) = ctx.plan, ctx.meta
        if u is None and all(r is None for r in rest):
            return (None,) * 9
        if any(r is not None for r in rest):
            fn = _alpha_expr(heads, ach, 1e-5, plan.dst, plan.N)
            with torch.enable_grad():
                leaves = [t.detach().requires_grad_(True) if t is not None else None for t in (Ya, ln_w, ln_b, alpha_dot)]
                gs = galpha.detach().requires_grad_(True)
                live = [t for t in leaves if t is not None]
                g1 = torch.autograd.grad(fn(*leaves), live, gs, create_graph=True)
                cots = (u,) + tuple(rest)
                cot = [c if c is not None else torch.zeros_like(g) for c, g in zip(cots, g1)]
                d = list(torch.autograd.grad(g1, live + [gs], cot, allow_unused=True))
            d_gs = d.pop()
            full = [d.pop(0) if t is not None else None for t in leaves]
            return full[0], full[1], full[2], full[3], None, d_gs, None, None, None
        E = Ya.shape[0]
        dev = Ya.device
        u = u.contiguous()
        R = torch.empty_like(alpha)
        d2Y = torch.empty_like(Ya)
        d_lnw = _small_zeros(ln_w.shape, dev) if ln_w is not None else None
        d_lnb = _small_zeros(ln_b.shape, dev) if ln_b is not None else None
        d_dot = _small_zeros(alpha_dot.shape, dev)
        d_galpha, d_alpha = torch.empty_like(alpha), torch.empty_like(alpha)
        _lib.call("eqv2_attn_alpha_bwd2", Ya.data_ptr(), heads * ach, _lib.ptr(ln_w), _lib.ptr(ln_b), alpha_dot.data_ptr(),
                  plan.rowptr_dst.data_ptr(), plan.perm_dst.data_ptr(), alpha.data_ptr(), galpha.data_ptr(),
                  dlogits.data_ptr(), u.data_ptr(), heads * ach, R.data_ptr(), d2Y.data_ptr(), heads * ach,
                  _lib.ptr(d_lnw), _lib.ptr(d_lnb), d_dot.data_ptr(), d_galpha.data_ptr(), d_alpha.data_ptr(), E, plan.N,
                  heads, ach, 1e-5, _lib.stream_ptr(), n_kernels=2)
        return d2Y, d_lnw, d_lnb, d_dot, d_alpha, d_galpha, None, None, None


def norm_groups(norm_type, lmax):
    """(ngroups, group_of_l, bw_l) for layer_norm.py:38-108 / :112-201 / :265-351."""
    if norm_type == "rms_norm_sh":
        return 1, [0] * (lmax + 1), [1.0 / ((2 * l + 1) * (lmax + 1)) for l in range(lmax + 1)]
    if norm_type == "layer_norm_sh":
        return (2 if lmax > 0 else 1), [0] + [1] * lmax, [1.0] + [1.0 / ((2 * l + 1) * lmax) for l in range(1, lmax + 1)]
    if norm_type == "layer_norm":
        return lmax + 1, list(range(lmax + 1)), [1.0 / (2 * l + 1) for l in range(lmax + 1)]
    raise ValueError(norm_type)


_norm_tables = {}     # (norm_type, lmax, device, dtype) -> (lk, bwk, gk): built once (no host->device copies in a hot loop / CUDA graph)


def _equiv_norm_tables(norm_type, lmax, dev, dtype):
    key = (norm_type, lmax, str(dev), dtype)
    if key not in _norm_tables:
        ng, gol, bw = norm_groups(norm_type, lmax)
        lk = torch.tensor([l for l in range(lmax + 1) for _ in range(2 * l + 1)], device=dev)
        _norm_tables[key] = (lk, torch.tensor(bw, dtype=dtype, device=dev)[lk], torch.tensor(gol, device=dev)[lk])
    return _norm_tables[key]


def _equiv_norm_expr(norm_type, lmax, eps):
    ng, gol, bw = norm_groups(norm_type, lmax)

    def fn(x, w, b):
        N, K, C = x.shape
        lk, bwk, gk = _equiv_norm_tables(norm_type, lmax, x.device, x.dtype)
        f = torch.cat([x[:, :1] - x[:, :1].mean(dim=2, keepdim=True), x[:, 1:]], dim=1)
        onehot = torch.nn.functional.one_hot(gk, ng).to(x.dtype)
        s = torch.einsum("nkc,k,kg->ng", f * f, bwk, onehot) / C
        inv = (s + eps).rsqrt()[:, gk]
        out = f * inv.unsqueeze(-1) * w[lk].unsqueeze(0)
        return torch.cat([out[:, :1] + b.view(1, 1, C), out[:, 1:]], dim=1)
    return fn


class EquivNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, norm_type, lmax, eps):
        _lib.check_device(x, w, b)
        assert x.is_contiguous() and w.is_contiguous() and b.is_contiguous()
        N, K, C = x.shape
        ng, gol, bw = norm_groups(norm_type, lmax)
        gol_c = (ctypes.c_int * len(gol))(*gol)
        bw_c = (ctypes.c_float * len(bw))(*bw)
        out = torch.empty_like(x)
        inv = torch.empty(N, ng, dtype=_F32, device=x.device)
        mean = torch.empty(N, dtype=_F32, device=x.device)
        slot = _absmax_slot(x.device)       # max |out|: operand bound of the gather/rotate kernel that consumes it
        _lib.call("eqv2_equiv_norm_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), inv.data_ptr(),
                  mean.data_ptr(), N, C, lmax, ng, ctypes.cast(gol_c, ctypes.c_void_p),
                  ctypes.cast(bw_c, ctypes.c_void_p), float(eps), _lib.ptr(slot), _lib.stream_ptr(),
                  work=(0.0, 8.0 * N * K * C))
        _register_absmax(out, slot)
        ctx.save_for_backward(x, w, b, inv, mean)
        ctx.meta = (norm_type, lmax, float(eps))
        return out

    @staticmethod
    def backward(ctx, go):
        x, w, b, inv, mean = ctx.saved_tensors
        norm_type, lmax, eps = ctx.meta
        if _recording(x, w, b, go):        # forces by autograd: the backward is itself differentiated (EquivNormBwdFn)
            gx, gw, gb = EquivNormBwdFn.apply(x, w, b, go.contiguous(), inv, mean, ctx.meta)
            return gx, gw, gb, None, None, None
        gx, gw, gb = _equiv_norm_bwd(x, w, go.contiguous(), inv, mean, norm_type, lmax)
        return gx, gw, gb, None, None, None


def _equiv_norm_bwd(x, w, go, inv, mean, norm_type, lmax):
    N, K, C = x.shape
    ng, gol, bw = norm_groups(norm_type, lmax)
    gol_c = (ctypes.c_int * len(gol))(*gol)
    bw_c = (ctypes.c_float * len(bw))(*bw)
    gx = torch.empty_like(x)
    gw = _small_zeros(w.shape, w.device)
    gb = _small_zeros((C,), x.device)
    _lib.call("eqv2_equiv_norm_bwd", x.data_ptr(), w.data_ptr(), go.data_ptr(), inv.data_ptr(), mean.data_ptr(),
              gx.data_ptr(), gw.data_ptr(), gb.data_ptr(), N, C, lmax, ng, ctypes.cast(gol_c, ctypes.c_void_p),
              ctypes.cast(bw_c, ctypes.c_void_p), _lib.stream_ptr(), work=(0.0, 12.0 * N * K * C))
    return gx, gw, gb


class EquivNormBwdFn(torch.autograd.Function):
    """First-order backward of EquivNormFn as a differentiable operator; its backward is the closed-form kernel
    `eqv2_equiv_norm_bwd2` for a cotangent of gx (the force loss).  Cotangents of the parameter gradients take the generic
    torch-expression route (`_second_order`-style)."""

    @staticmethod
    def forward(ctx, x, w, b, go, inv, mean, meta):
        norm_type, lmax, eps = meta
        gx, gw, gb = _equiv_norm_bwd(x, w, go, inv, mean, norm_type, lmax)
        ctx.save_for_backward(x, w, b, go, inv, mean)
        ctx.meta = meta
        ctx.set_materialize_grads(False)
        return gx, gw, gb

    @staticmethod
    def backward(ctx, u, uw, ub):
        x, w, b, go, inv, mean = ctx.saved_tensors
        norm_type, lmax, eps = ctx.meta
        if u is None and uw is None and ub is None:
            return None, None, None, None, None, None, None
        if uw is not None or ub is not None:
            fn = _equiv_norm_expr(norm_type, lmax, eps)
            with torch.enable_grad():
                xs, ws, bs, gs = [t.detach().requires_grad_(True) for t in (x, w, b, go)]
                g1 = torch.autograd.grad(fn(xs, ws, bs), [xs, ws, bs], gs, create_graph=True)
                cot = [c if c is not None else torch.zeros_like(g) for c, g in zip((u, uw, ub), g1)]
                d = torch.autograd.grad(g1, [xs, ws, bs, gs], cot, allow_unused=True)
            return d[0], d[1], d[2], d[3], None, None, None
        N, K, C = x.shape
        ng, gol, bw = norm_groups(norm_type, lmax)
        gol_c = (ctypes.c_int * len(gol))(*gol)
        bw_c = (ctypes.c_float * len(bw))(*bw)
        u = u.contiguous()
        d2x, dgo = torch.empty_like(x), torch.empty_like(go)
        dw = _small_zeros(w.shape, w.device)
        _lib.call("eqv2_equiv_norm_bwd2", x.data_ptr(), w.data_ptr(), go.data_ptr(), inv.data_ptr(), mean.data_ptr(),
                  u.data_ptr(), d2x.data_ptr(), dgo.data_ptr(), dw.data_ptr(), N, C, lmax, ng,
                  ctypes.cast(gol_c, ctypes.c_void_p), ctypes.cast(bw_c, ctypes.c_void_p), _lib.stream_ptr(),
                  work=(0.0, 20.0 * N * K * C))
        return d2x, dw, None, dgo, None, None, None


def _ln_silu_bwd(x, w, b, gy, eps):
    rows, width = x.shape
    gx = torch.empty_like(x)
    gw = _small_zeros(w.shape, w.device)
    gb = _small_zeros(b.shape, b.device)
    _lib.call("eqv2_ln_silu_bwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), gy.data_ptr(), gx.data_ptr(),
              gw.data_ptr(), gb.data_ptr(), rows, width, eps, _lib.stream_ptr(), work=(0.0, 12.0 * rows * width))
    return gx, gw, gb


class LnSiluFn(torch.autograd.Function):
    """SiLU(LayerNorm(x)) of the radial MLP (radial_function.py:21-22)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        _lib.check_device(x, w, b)
        assert x.is_contiguous() and w.is_contiguous() and b.is_contiguous()
        rows, width = x.shape
        y = torch.empty_like(x)
        _lib.call("eqv2_ln_silu_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, width, float(eps),
                  _lib.stream_ptr(), work=(0.0, 8.0 * rows * width))
        ctx.save_for_backward(x, w, b)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, b = ctx.saved_tensors
        if _recording(x, w, b, gy):        # forces by autograd: the backward is itself differentiated (LnSiluBwdFn)
            gx, gw, gb = LnSiluBwdFn.apply(x, w, b, gy.contiguous(), ctx.eps)
            return gx, gw, gb, None
        gx, gw, gb = _ln_silu_bwd(x, w, b, gy.contiguous(), ctx.eps)
        return gx, gw, gb, None


class LnSiluBwdFn(torch.autograd.Function):
    """The first-order backward of LnSiluFn as a differentiable operator: (x, w, b, gy) -> (gx, gw, gb).  Its own
    backward is the closed-form kernel `eqv2_ln_silu_bwd2` for a cotangent of gx -- the case of the force loss, where
    the first backward only feeds the position gradient.  Cotangents of gw / gb (somebody differentiating the PARAMETER
    gradients again) take the generic route: the operator re-expressed with torch primitives (`_second_order`)."""

    @staticmethod
    def forward(ctx, x, w, b, gy, eps):
        gx, gw, gb = _ln_silu_bwd(x, w, b, gy, eps)
        ctx.save_for_backward(x, w, b, gy)
        ctx.eps = eps
        ctx.set_materialize_grads(False)      # cotangents of unused outputs (gw, gb in the force pass) arrive as None
        return gx, gw, gb

    @staticmethod
    def backward(ctx, u, uw, ub):
        x, w, b, gy = ctx.saved_tensors
        eps = ctx.eps
        if u is None and uw is None and ub is None:
            return None, None, None, None, None
        if uw is not None or ub is not None:
            Fn = torch.nn.functional
            with torch.enable_grad():
                xs, ws, bs, gs = [t.detach().requires_grad_(True) for t in (x, w, b, gy)]
                y = Fn.silu(Fn.layer_norm(xs, xs.shape[-1:], ws, bs, eps))
                g1 = torch.autograd.grad(y, [xs, ws, bs], gs, create_graph=True)
                cot = [c if c is not None else torch.zeros_like(g) for c, g in zip((u, uw, ub), g1)]
                d = torch.autograd.grad(g1, [xs, ws, bs, gs], cot, allow_unused=True)
            return d[0], d[1], d[2], d[3], None
        rows, width = x.shape
        u = u.contiguous()
        dx, dgy = torch.empty_like(x), torch.empty_like(gy)
        dw = _small_zeros(w.shape, w.device)
        db = _small_zeros(b.shape, b.device)
        _lib.call("eqv2_ln_silu_bwd2", x.data_ptr(), w.data_ptr(), b.data_ptr(), gy.data_ptr(), u.data_ptr(), dx.data_ptr(),
                  dgy.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, width, eps, _lib.stream_ptr(),
                  work=(0.0, 20.0 * rows * width))
        return dx, dw, db, dgy, None


class RbfFn(torch.autograd.Function):
    """GaussianSmearing (equiformerv2_oc20.py:43-60): exp(coeff * (d - offset_k)^2)."""

    @staticmethod
    def forward(ctx, d, offset, coeff):
        _lib.check_device(d, offset)
        assert d.dim() == 1 and d.is_contiguous() and offset.is_contiguous()
        R = offset.shape[0]
        out = torch.empty(d.shape[0], R, dtype=_F32, device=d.device)
        _lib.call("eqv2_rbf_fwd", d.data_ptr(), out.data_ptr(), d.shape[0], R, offset.data_ptr(), float(coeff),
                  _lib.stream_ptr(), work=(0.0, 4.0 * d.shape[0] * (R + 1)))
        ctx.save_for_backward(d, offset)
        ctx.coeff = float(coeff)
        return out

    @staticmethod
    def backward(ctx, go):
        d, offset = ctx.saved_tensors
        if _recording(d, go):              # forces by autograd: differentiable backward (RbfBwdFn)
            return RbfBwdFn.apply(d, go.contiguous(), offset, ctx.coeff), None, None
        return _rbf_bwd(d, go.contiguous(), offset, ctx.coeff), None, None


def _rbf_bwd(d, go, offset, coeff):
    gd = torch.empty_like(d)
    _lib.call("eqv2_rbf_bwd", d.data_ptr(), go.data_ptr(), gd.data_ptr(), d.shape[0], offset.shape[0],
              offset.data_ptr(), coeff, _lib.stream_ptr())
    return gd


class RbfBwdFn(torch.autograd.Function):
    """dd = sum_k go_k d/dd rbf_k(d) as a differentiable operator of (d, go); its backward is `eqv2_rbf_bwd2`."""

    @staticmethod
    def forward(ctx, d, go, offset, coeff):
        ctx.save_for_backward(d, go, offset)
        ctx.coeff = coeff
        return _rbf_bwd(d, go, offset, coeff)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        d, go, offset = ctx.saved_tensors
        u = u.contiguous()
        dgo, d2d = torch.empty_like(go), torch.empty_like(d)
        _lib.call("eqv2_rbf_bwd2", d.data_ptr(), go.data_ptr(), u.data_ptr(), dgo.data_ptr(), d2d.data_ptr(), d.shape[0],
                  offset.shape[0], offset.data_ptr(), ctx.coeff, _lib.stream_ptr())
        return d2d, dgo, None, None


# ----------------------------------------------------------------------------------------------
# GaussianSmearing + first radial-MLP layer, fused (north_star piece 2)
# ----------------------------------------------------------------------------------------------
RBF_BAND_TOL = 1e-12          # neglected basis functions are below this (relative to the largest one, which is ~1)
_FEATURES["rbf_linear"] = "rbf_linear" not in _os.environ.get("EQV2_DISABLE", "").split(",")


class RbfSource:
    """What GaussianSmearing attaches to the [E, R] tensor it returns (`tensor._eqv2_rbf`): the raw distances and the
    basis description, so that a consumer that only needs W . rbf(d) can evaluate it from d without reading the tensor."""
    __slots__ = ("dist", "offset", "coeff", "start", "delta", "band", "_bins")

    def __init__(self, dist, offset, coeff, start, delta):
        self.dist, self.offset, self.coeff, self.start, self.delta = dist, offset, float(coeff), float(start), float(delta)
        # exp(coeff (j delta)^2) < tol  <=>  j > sqrt(ln(tol) / (coeff delta^2))
        self.band = int(math.ceil(math.sqrt(math.log(RBF_BAND_TOL) / (self.coeff * self.delta ** 2))))
        self._bins = None

    def bins(self):
        """Plan of the banded weight gradient (eqv2_rbf_linear_wgrad), built once per graph with device-side torch ops (no
        read-back) and shared by every block: edges sorted by nearest basis index; per chunk of sorted edges the range of
        basis functions it can touch and the first row of its partial sums; per basis function the chunks that touch it.
        -> (perm, chunk_k [nchunks, 2], chunk_base [nchunks], k_chunks [R, 2], partial_rows)"""
        if self._bins is None:
            R, E = int(self.offset.shape[0]), int(self.dist.shape[0])
            CH = int(_lib.lib().eqv2_rbf_linear_chunk())
            dev = self.dist.device
            k0 = torch.round((self.dist - self.start) / self.delta).clamp_(0, R - 1).long()
            k0s, perm = torch.sort(k0, stable=True)
            nchunks = max(1, -(-E // CH))
            first_i = torch.arange(nchunks, device=dev) * CH
            last_i = torch.clamp(first_i + CH - 1, max=max(E - 1, 0))
            if E > 0:
                kmin = (k0s[first_i] - self.band).clamp_(min=0)
                kmax = (k0s[last_i] + self.band).clamp_(max=R - 1)
            else:
                kmin = torch.zeros(1, dtype=torch.long, device=dev)
                kmax = torch.full((1,), -1, dtype=torch.long, device=dev)
            nk = kmax - kmin + 1
            base = torch.cumsum(nk, 0) - nk
            ks = torch.arange(R, device=dev)
            clo = torch.searchsorted(kmax, ks, right=False)
            chi = torch.searchsorted(kmin, ks, right=True) - 1
            self._bins = (perm.to(torch.int32), torch.stack([kmin, kmax], 1).to(torch.int32).contiguous(),
                          base.to(torch.int32), torch.stack([clo, chi], 1).to(torch.int32).contiguous(),
                          R + nchunks * (2 * self.band + 1))           # upper bound of sum(nk): no device read-back
        return self._bins


class FusedEdgeFeatures:
    """x_edge = [rbf(d) | E_src[Z_src] | E_dst[Z_dst]] (transformer_block.py:241-248) as a DESCRIPTION: the radial MLP's
    first layer evaluates W1 x_edge + b1 from it directly (`rbf_linear`), the [E, R + 2 Ce] matrix is never formed."""

    def __init__(self, src, src_w, dst_w, zs, csr_s, zd, csr_d):
        self.src, self.src_w, self.dst_w = src, src_w, dst_w
        self.zs, self.csr_s, self.zd, self.csr_d = zs, csr_s, zd, csr_d

    @property
    def width(self):
        return int(self.src.offset.shape[0]) + (2 * int(self.src_w.shape[1]) if self.src_w is not None else 0)

    def dense(self):
        """The [E, R + 2 Ce] matrix after all (consumers the fused kernel does not cover)."""
        rbf_t = RbfFn.apply(self.src.dist, self.src.offset.contiguous(), self.src.coeff)
        if self.src_w is None:
            return rbf_t
        return torch.cat((rbf_t, embed_rows(self.src_w, self.zs, self.csr_s), embed_rows(self.dst_w, self.zd, self.csr_d)),
                         dim=1)


def rbf_source_of(t):
    """The RbfSource behind an rbf tensor if the fused first layer may use it: feature on, CUDA library has the kernels,
    distances carry no gradient (first-order step)."""
    src = getattr(t, "_eqv2_rbf", None)
    if src is None or not _FEATURES["rbf_linear"] or not hasattr(_lib.lib(), "eqv2_rbf_linear_fwd"):
        return None
    if torch.is_grad_enabled() and src.dist.requires_grad:
        return None
    return src


class RbfLinearFn(torch.autograd.Function):
    """h = Wt^T rbf(d) + Ts[zs] + Td[zd] + b over the (2 band + 1) nearest basis functions.  Backward: banded weight
    gradient in bin order, table gradients as deterministic segmented column sums, bias gradient as a column sum."""

    @staticmethod
    def forward(ctx, Wt, Ts, Td, bias, feat):
        src = feat.src
        _lib.check_device(Wt, Ts, Td, bias, src.dist)
        assert Wt.is_contiguous() and (Ts is None or (Ts.is_contiguous() and Td.is_contiguous()))
        E, (R, H) = int(src.dist.shape[0]), Wt.shape
        out = torch.empty(E, H, dtype=_F32, device=Wt.device)
        _lib.call("eqv2_rbf_linear_fwd", src.dist.data_ptr(), src.offset.data_ptr(), Wt.data_ptr(), _lib.ptr(Ts),
                  _lib.ptr(Td), _lib.ptr(feat.zs) if Ts is not None else None, _lib.ptr(feat.zd) if Ts is not None else None,
                  _lib.ptr(bias), out.data_ptr(), E, int(R), int(H), src.start, src.delta, src.coeff, src.band,
                  _lib.stream_ptr(), work=(2.0 * E * H * (2 * src.band + 1), 4.0 * E * (H + 1)))
        ctx.feat, ctx.shape = feat, (int(R), int(H))
        ctx.has = (Ts is not None, bias is not None)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gh):
        feat, (R, H) = ctx.feat, ctx.shape
        src = feat.src
        gh = gh.contiguous()
        gWt = gTs = gTd = gb = None
        if ctx.needs_input_grad[0]:
            perm, chunk_k, chunk_base, k_chunks, rows = src.bins()
            gWt = torch.empty(R, H, dtype=_F32, device=gh.device)
            partial = torch.empty(rows, H, dtype=_F32, device=gh.device)
            _lib.call("eqv2_rbf_linear_wgrad", src.dist.data_ptr(), src.offset.data_ptr(), perm.data_ptr(),
                      chunk_k.data_ptr(), chunk_base.data_ptr(), k_chunks.data_ptr(), gh.data_ptr(), partial.data_ptr(),
                      gWt.data_ptr(), int(gh.shape[0]), R, H, src.coeff, _lib.stream_ptr(), n_kernels=2,
                      work=(2.0 * gh.shape[0] * H * (2 * src.band + 1), 4.0 * gh.shape[0] * H))
        if ctx.has[0]:
            V = int(feat.src_w.shape[0])
            S = max(1, min(256, int(gh.shape[0]) // 32))
            if ctx.needs_input_grad[1]:
                gTs = _seg_colsum(gh, H, 0, gh.shape[0], H, V, feat.csr_s[1], feat.csr_s[0], S)
            if ctx.needs_input_grad[2]:
                gTd = _seg_colsum(gh, H, 0, gh.shape[0], H, V, feat.csr_d[1], feat.csr_d[0], S)
        if ctx.has[1] and ctx.needs_input_grad[3]:
            gb = colsum(gh, 0, H)
        return gWt, gTs, gTd, gb, None


def rbf_linear(feat, W1, b1):
    """First layer of the radial MLP on a FusedEdgeFeatures description: the weight slices are taken apart with ordinary
    (differentiable) torch ops -- [H, R] -> transposed [R, H] for coalesced row reads; the embedding tables are multiplied
    by their slices once per element type ([V, Ce] x [Ce, H], V ~ 90) instead of once per edge."""
    R = int(feat.src.offset.shape[0])
    Wt = W1[:, :R].t().contiguous()
    Ts = Td = None
    if feat.src_w is not None:
        Ce = int(feat.src_w.shape[1])
        Ts = torch.mm(feat.src_w, W1[:, R:R + Ce].t())
        Td = torch.mm(feat.dst_w, W1[:, R + Ce:R + 2 * Ce].t())
    return RbfLinearFn.apply(Wt, Ts, Td, b1, feat)


# ----------------------------------------------------------------------------------------------
# GATA / HTR per-edge operators (configs 4-5), differentiable twice through kernels only (csrc/gata.cu)
# ----------------------------------------------------------------------------------------------
class HtrInnerFn(torch.autograd.Function):
    """T(q, k; r) [E, H]: the sum over degrees of <reject(q^l, r^l), reject(k^l, -r^l)> / (2l+1)  (HTR, activation.py:
    166-264).  Bilinear; its gradient map is HtrGradFn, and the pair is closed under differentiation."""

    @staticmethod
    def forward(ctx, q, k, rl, lmax):
        _lib.check_device(q, k, rl)
        assert q.is_contiguous() and k.is_contiguous() and rl.is_contiguous() and q.shape == k.shape
        E, M, H = q.shape
        assert M == (lmax + 1) ** 2 - 1 and rl.shape == (E, M)
        out = torch.empty(E, H, dtype=_F32, device=q.device)
        _lib.call("eqv2_htr_inner", q.data_ptr(), k.data_ptr(), rl.data_ptr(), out.data_ptr(), E, H, lmax, _lib.stream_ptr(),
                  work=(8.0 * E * M * H, 4.0 * E * (2 * M * H + H)))
        ctx.save_for_backward(q, k, rl)
        ctx.lmax = lmax
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, rl = ctx.saved_tensors
        g = g.contiguous()
        dq = HtrGradFn.apply(g, k, rl, ctx.lmax) if ctx.needs_input_grad[0] else None
        dk = HtrGradFn.apply(g, q, rl, ctx.lmax) if ctx.needs_input_grad[1] else None
        return dq, dk, None, None


class HtrGradFn(torch.autograd.Function):
    """G(g, b; r) [E, M, H] = g / (2l+1) * (b - (2 - |r^l|^2)(b^l.r^l) r): <u, G(g, b)> = g . T(u, b)."""

    @staticmethod
    def forward(ctx, g, b, rl, lmax):
        assert g.is_contiguous() and b.is_contiguous()
        E, M, H = b.shape
        out = torch.empty_like(b)
        _lib.call("eqv2_htr_grad", g.data_ptr(), b.data_ptr(), rl.data_ptr(), out.data_ptr(), E, H, lmax, _lib.stream_ptr(),
                  work=(6.0 * E * M * H, 4.0 * E * (2 * M * H + H)))
        ctx.save_for_backward(g, b, rl)
        ctx.lmax = lmax
        return out

    @staticmethod
    def backward(ctx, u):
        g, b, rl = ctx.saved_tensors
        u = u.contiguous()
        dg = HtrInnerFn.apply(u, b, rl, ctx.lmax) if ctx.needs_input_grad[0] else None
        db = HtrGradFn.apply(g, u, rl, ctx.lmax) if ctx.needs_input_grad[1] else None
        return dg, db, None, None


def htr_inner(q, k, rl, lmax):
    return HtrInnerFn.apply(q.contiguous(), k.contiguous(), rl.contiguous(), lmax)


def gata_rows(lmax, mmax):
    return 1 + sum(min(2 * l + 1, 2 * mmax + 1) for l in range(1, lmax + 1))


class GataValueFn(torch.autograd.Function):
    """GATAValueActivation's output assembly (activation.py:370-414): SiLU on the scalar row, o_d^l r^l + o_t^l Xp^l on the
    first min(2l+1, 2 mmax+1) rows of every degree.  One kernel per derivative order (fwd / bwd / bwd-of-bwd)."""

    @staticmethod
    def forward(ctx, comb, Xp, rl, lmax, mmax):
        _lib.check_device(comb, Xp, rl)
        assert comb.is_contiguous() and Xp.is_contiguous() and rl.is_contiguous()
        E, M, H = Xp.shape
        assert comb.shape == (E, (1 + 2 * lmax) * H) and rl.shape == (E, M) and M == (lmax + 1) ** 2 - 1
        Kr = gata_rows(lmax, mmax)
        out = torch.empty(E, Kr, H, dtype=_F32, device=comb.device)
        _lib.call("eqv2_gata_value_fwd", comb.data_ptr(), Xp.data_ptr(), rl.data_ptr(), out.data_ptr(), E, H, lmax, mmax, Kr,
                  _lib.stream_ptr(), work=(4.0 * E * Kr * H, 4.0 * E * H * (1 + 2 * lmax + M + Kr)))
        ctx.save_for_backward(comb, Xp, rl)
        ctx.lm = (lmax, mmax)
        return out

    @staticmethod
    def backward(ctx, g):
        comb, Xp, rl = ctx.saved_tensors
        dcomb, dXp = GataValueBwdFn.apply(comb, Xp, rl, g.contiguous(), *ctx.lm)
        return dcomb, dXp, None, None, None


class GataValueBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, comb, Xp, rl, g, lmax, mmax):
        assert g.is_contiguous()
        E, M, H = Xp.shape
        Kr = gata_rows(lmax, mmax)
        dcomb, dXp = torch.empty_like(comb), torch.empty_like(Xp)
        _lib.call("eqv2_gata_value_bwd", comb.data_ptr(), Xp.data_ptr(), rl.data_ptr(), g.data_ptr(), dcomb.data_ptr(),
                  dXp.data_ptr(), E, H, lmax, mmax, Kr, _lib.stream_ptr(),
                  work=(4.0 * E * Kr * H, 4.0 * E * H * (2 * (1 + 2 * lmax) + 2 * M + Kr)))
        ctx.save_for_backward(comb, Xp, rl, g)
        ctx.lm = (lmax, mmax)
        return dcomb, dXp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, v):
        comb, Xp, rl, g = ctx.saved_tensors
        lmax, mmax = ctx.lm
        E, M, H = Xp.shape
        Kr = gata_rows(lmax, mmax)
        u = u.contiguous() if u is not None else None
        v = v.contiguous() if v is not None else None
        dg, d2comb, d2Xp = torch.empty_like(g), torch.empty_like(comb), torch.empty_like(Xp)
        _lib.call("eqv2_gata_value_bwd2", comb.data_ptr(), Xp.data_ptr(), rl.data_ptr(), g.data_ptr(), _lib.ptr(u),
                  _lib.ptr(v), dg.data_ptr(), d2comb.data_ptr(), d2Xp.data_ptr(), E, H, lmax, mmax, Kr, _lib.stream_ptr(),
                  work=(8.0 * E * Kr * H, 4.0 * E * H * (3 * (1 + 2 * lmax) + 3 * M + 2 * Kr)))
        return d2comb, d2Xp, None, dg, None, None


# ----------------------------------------------------------------------------------------------
# all-pairs attention core of the global-attention classes (csrc/pair_attn.cu)
# ----------------------------------------------------------------------------------------------
class PairLayout:
    """Ragged [P, H] layout of the per-structure attention maps: row i owns gcount[i] contiguous entries from rowptr[i]
    (see include/eqv2_b200.h).  Built from the host-side structure sizes: no device read-back."""

    _cache = {}

    def __init__(self, counts, device):
        counts = [int(n) for n in counts]
        self.counts, self.N = counts, sum(counts)
        self.max_count = max(counts) if counts else 0
        self.P = sum(n * n for n in counts)
        gs, gc, rp = [], [], [0]
        start = 0
        for n in counts:
            for _ in range(n):
                gs.append(start)
                gc.append(n)
                rp.append(rp[-1] + n)
            start += n
        self.gstart = torch.tensor(gs, dtype=torch.int32, device=device)
        self.gcount = torch.tensor(gc, dtype=torch.int32, device=device)
        self.rowptr = torch.tensor(rp, dtype=torch.int64, device=device)

    @classmethod
    def get(cls, counts, device):
        key = (tuple(int(n) for n in counts), str(device))
        hit = cls._cache.get(key)
        if hit is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            hit = cls._cache[key] = cls(counts, device)
        return hit

    def from_blocks(self, blocks):
        """[H, n_g, n_g] per structure -> [P, H]."""
        return torch.cat([b.permute(1, 2, 0).reshape(-1, b.shape[0]) for b in blocks], dim=0)


def _pair_scores(a, b, lay, scale):
    N, M, H, D = a.shape
    S = torch.empty(lay.P, H, dtype=_F32, device=a.device)
    _lib.call("eqv2_pair_scores", a.data_ptr(), b.data_ptr(), lay.gstart.data_ptr(), lay.gcount.data_ptr(),
              lay.rowptr.data_ptr(), S.data_ptr(), N, M, H, D, float(scale), _lib.stream_ptr(),
              work=(2.0 * lay.P * M * H * D, 4.0 * (2 * N * M * H * D + lay.P * H)))
    return S


def _pair_mix(W, b, lay, transpose):
    N, M, H, D = b.shape
    out = torch.empty_like(b)
    _lib.call("eqv2_pair_mix", W.data_ptr(), b.data_ptr(), lay.gstart.data_ptr(), lay.gcount.data_ptr(),
              lay.rowptr.data_ptr(), out.data_ptr(), N, M, H, D, lay.max_count, 1 if transpose else 0, _lib.stream_ptr(),
              work=(2.0 * lay.P * M * H * D, 4.0 * (2 * N * M * H * D + lay.P * H)))
    return out


class PairScoresFn(torch.autograd.Function):
    """S[pair(i,j), h] = scale <a_i, b_j>_h  (activation.py:1533 `einsum('bihd,bjhd->bhij')` per structure)."""

    @staticmethod
    def forward(ctx, a, b, lay, scale):
        _lib.check_device(a, b)
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        ctx.lay, ctx.scale = lay, scale
        return _pair_scores(a, b, lay, scale)

    @staticmethod
    def backward(ctx, gS):
        a, b = ctx.saved_tensors
        ga = gb = None
        if gS is not None:
            gS = gS.contiguous()
            if ctx.needs_input_grad[0]:
                ga = PairMixFn.apply(gS, b, ctx.lay, False) * ctx.scale
            if ctx.needs_input_grad[1]:
                gb = PairMixFn.apply(gS, a, ctx.lay, True) * ctx.scale
        return ga, gb, None, None


class PairMixFn(torch.autograd.Function):
    """out_i = sum_j W[pair(i,j)] b_j  (activation.py:1545 `einsum('bhij,bjmhd->bimhd')`); transpose: out_j = sum_i W b_i."""

    @staticmethod
    def forward(ctx, W, b, lay, transpose):
        _lib.check_device(W, b)
        W, b = W.contiguous(), b.contiguous()
        ctx.save_for_backward(W, b)
        ctx.lay, ctx.transpose = lay, transpose
        return _pair_mix(W, b, lay, transpose)

    @staticmethod
    def backward(ctx, gout):
        W, b = ctx.saved_tensors
        gW = gb = None
        if gout is not None:
            gout = gout.contiguous()
            if ctx.needs_input_grad[0]:
                gW = PairScoresFn.apply(b, gout, ctx.lay, 1.0) if ctx.transpose else PairScoresFn.apply(gout, b, ctx.lay, 1.0)
            if ctx.needs_input_grad[1]:
                gb = PairMixFn.apply(W, gout, ctx.lay, not ctx.transpose)
        return gW, gb, None, None


class PairSoftmaxFn(torch.autograd.Function):
    """softmax over the atoms of the query's structure (the reference's masked / padded softmax, activation.py:1541)."""

    @staticmethod
    def forward(ctx, S, lay):
        _lib.check_device(S)
        S = S.contiguous()
        Pw = torch.empty_like(S)
        _lib.call("eqv2_pair_softmax_fwd", S.data_ptr(), lay.gcount.data_ptr(), lay.rowptr.data_ptr(), Pw.data_ptr(),
                  lay.N, S.shape[1], _lib.stream_ptr(), work=(0.0, 8.0 * S.numel()))
        ctx.save_for_backward(Pw)
        ctx.lay = lay
        return Pw

    @staticmethod
    def backward(ctx, gP):
        (Pw,) = ctx.saved_tensors
        return PairSoftmaxBwdFn.apply(Pw, gP, ctx.lay), None


class PairSoftmaxBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Pw, gP, lay):
        gP = gP.contiguous()
        gS = torch.empty_like(Pw)
        _lib.call("eqv2_pair_softmax_bwd", Pw.data_ptr(), gP.data_ptr(), lay.gcount.data_ptr(), lay.rowptr.data_ptr(),
                  gS.data_ptr(), lay.N, Pw.shape[1], _lib.stream_ptr(), work=(0.0, 12.0 * Pw.numel()))
        ctx.save_for_backward(Pw, gP)
        ctx.lay = lay
        return gS

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        Pw, gP = ctx.saved_tensors
        lay = ctx.lay
        u = u.contiguous()
        dP, dgP = torch.empty_like(Pw), torch.empty_like(Pw)
        _lib.call("eqv2_pair_softmax_bwd2", Pw.data_ptr(), gP.data_ptr(), u.data_ptr(), lay.gcount.data_ptr(),
                  lay.rowptr.data_ptr(), dP.data_ptr(), dgP.data_ptr(), lay.N, Pw.shape[1], _lib.stream_ptr(),
                  work=(0.0, 20.0 * Pw.numel()))
        return dP, dgP, None


def pair_attention(q, k, values, counts, scale, dropout=None, bias_blocks=None):
    """Per-structure multi-head attention for all structures at once: q, k [N, H, D]; values: list of [N, m, H, D] mixed
    with the SAME weights; bias_blocks: list of [H, n_g, n_g] additive logit biases or None.  -> list of [N, m, H*D]."""
    N, H, D = q.shape
    lay = PairLayout.get(counts, q.device)
    S = PairScoresFn.apply(q.reshape(N, 1, H, D), k.reshape(N, 1, H, D), lay, scale)
    if bias_blocks is not None:
        S = S + lay.from_blocks(bias_blocks)
    Pw = PairSoftmaxFn.apply(S, lay)
    if dropout is not None:
        Pw = dropout(Pw)
    sizes = [v.shape[1] for v in values]
    V = values[0] if len(values) == 1 else torch.cat(values, dim=1)
    out = PairMixFn.apply(Pw, V, lay, False)
    return [o.reshape(N, m, H * D) for o, m in zip(torch.split(out, sizes, dim=1), sizes)]


def pair_attention_available(H, D):
    return D % 4 == 0 and ((D // 4) & (D // 4 - 1)) == 0 and D // 4 <= 32


class DropPathScaleFn(torch.autograd.Function):
    """x * scale[batch], scale = floor(keep + u) / keep (GraphDropPath, drop.py:49-68): one kernel instead of the
    reference's seven element-wise launches; linear in x, so the backward is the same kernel on the gradient."""

    @staticmethod
    def forward(ctx, x, u, batch, keep):
        _lib.check_device(x, u, batch)
        x = x.contiguous()
        out = torch.empty_like(x)
        N = int(x.shape[0])
        _lib.call("eqv2_drop_path_scale", x.data_ptr(), u.data_ptr(), batch.data_ptr(), float(keep), out.data_ptr(), N,
                  x.numel() // max(N, 1), _lib.stream_ptr(), work=(0.0, 8.0 * x.numel()))
        ctx.save_for_backward(u, batch)
        ctx.keep = keep
        return out

    @staticmethod
    def backward(ctx, g):
        u, batch = ctx.saved_tensors
        return DropPathScaleFn.apply(g, u, batch, ctx.keep), None, None, None


def drop_path_scale(x, u, batch, keep):
    return DropPathScaleFn.apply(x, u.reshape(-1).contiguous(), batch.contiguous(), keep)


def gata_value(comb, Xp, rl, lmax, mmax):
    return GataValueFn.apply(comb.contiguous(), Xp.contiguous(), rl.contiguous(), lmax, mmax)


def gather_rotate(x, rad, plan, wig, lmax, mmax):
    return GatherRotateFn.apply(x.contiguous(), rad.contiguous() if rad is not None else None, plan, wig, lmax, mmax)


def rotinv_reduce(val, alpha, plan, wig, lmax, mmax, rows_used, heads, scale):
    return RotInvReduceFn.apply(val.contiguous(), alpha.contiguous() if alpha is not None else None, plan, wig, lmax, mmax,
                                rows_used, heads, scale)


def s2_act(x, gate, mats):
    return S2ActFn.apply(x.contiguous(), gate.contiguous() if gate is not None else None, mats)


def attn_alpha(Ya, ln_w, ln_b, alpha_dot, plan, heads, ach):
    return AttnAlphaFn.apply(Ya.contiguous(), ln_w, ln_b, alpha_dot.contiguous(), plan, heads, ach)


def edge_act_alpha(Y, ln_w, ln_b, alpha_dot, plan, mats, heads, ach, H, z_planes=False):
    """z_planes: Z is written as the operand planes of the second convolution; the returned Z is a stand-in that only
    `conv_rotinv_reduce` can consume (see s2_planes_available)."""
    Y = Y.contiguous()
    Z, alpha = EdgeActAlphaFn.apply(Y, ln_w, ln_b, alpha_dot.contiguous(), plan, mats, heads, ach, H, z_planes)
    if z_planes:
        Z._eqv2_planes, EdgeActAlphaFn.last_planes = EdgeActAlphaFn.last_planes, None
    return Z, alpha


def s2_planes_available(Y, mats, H, width, Cv, heads):
    """Can the S2 activation write the second convolution's A operand as planes?  f16 engine, the separable kernel, whole
    rows per 128-thread tile, the maximum of Y known (GEMM epilogue), and the conditions under which ConvRotInvReduceFn keeps
    Z as planes only."""
    return (_FEATURES["planes"] and _FEATURES["s2_planes"] and _GEMM_MODE["mode"] in ("f16x3", "f16") and Y.is_cuda
            and mats.factors is not None and 128 % H == 0 and H % 8 == 0 and (mats.Kr * H) % 8 == 0
            and width % 8 == 0 and Cv % 32 == 0 and heads <= Cv and _known_absmax(Y.contiguous()) is not None
            and hasattr(_lib.lib(), "eqv2_s2sep_fwd_planes") and hasattr(_lib.lib(), "eqv2_rotinv_reduce_bwd_planes"))


def equiv_norm(x, w, b, norm_type, lmax, eps):
    return EquivNormFn.apply(x.contiguous(), w.contiguous(), b.contiguous(), norm_type, lmax, eps)


def ln_silu(x, w, b, eps):
    return LnSiluFn.apply(x.contiguous(), w.contiguous(), b.contiguous(), eps)


def rbf(d, offset, coeff, start=None, delta=None):
    """GaussianSmearing.  With `start` / `delta` (uniformly spaced offsets) the result also carries an RbfSource, which
    lets the radial MLP's first layer skip the [E, R] tensor (rbf_linear)."""
    d = d.reshape(-1).contiguous()
    out = RbfFn.apply(d, offset.contiguous(), coeff)
    if start is not None and delta is not None and delta > 0:
        out._eqv2_rbf = RbfSource(d.detach() if not d.requires_grad else d, offset, coeff, start, delta)
    return out
