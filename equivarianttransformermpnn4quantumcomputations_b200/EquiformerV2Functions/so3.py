"""SO(3) containers and helpers with the reference's names and state_dict keys
(reference so3.py): CoefficientMappingModule (:45-199), SO3_Embedding (:203-487),
SO3_Rotation (:490-545), SO3_Grid (:552-646), SO3_LinearV2 (:698-743).

What changes underneath:
  * SO3_Rotation.set_wigner launches `eqv2_wigner_from_rot` and keeps the Wigner-D matrices
    block-diagonal ([E, sum (2l+1)^2]); the dense [E,K,K] tensors of the reference
    (`.wigner`, `.wigner_inv`) are materialised lazily only if somebody reads them.
  * SO3_Grid builds its matrices from first principles (no e3nn) and hands padded copies to the
    fused S2-activation kernel.
  * SO3_LinearV2 runs one grouped GEMM (a group per degree l) instead of an expanded einsum.
The SO3_Embedding helper methods keep their reference semantics for external callers; the
fused blocks in transformer_block.py do not go through them.
"""
import math

import torch
import torch.nn as nn

from .. import _so3_math, ops


class CoefficientMappingModule(nn.Module):
    """Index bookkeeping between degrees l and orders m; buffers as in the reference
    (`l_harmonic`, `m_harmonic`, `m_complex`, `res_size`, `to_m`, `m_size`)."""

    def __init__(self, lmax_list, mmax_list):
        super().__init__()
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list
        self.num_resolutions = len(lmax_list)
        self.device = "cpu"

        ls, ms, sizes = [], [], []
        for lmax, mmax in zip(lmax_list, mmax_list):
            before = len(ls)
            for l in range(lmax + 1):
                mm = min(mmax, l)
                for m in range(-mm, mm + 1):
                    ls.append(l)
                    ms.append(m)
            sizes.append(len(ls) - before)
        l_harmonic = torch.tensor(ls, dtype=torch.long)
        m_complex = torch.tensor(ms, dtype=torch.long)
        m_harmonic = m_complex.abs()
        n = len(ls)
        top_m = max(mmax_list)
        to_m = torch.zeros(n, n)
        m_size = torch.zeros(top_m + 1, dtype=torch.long)
        row = 0
        for m in range(top_m + 1):
            plus, minus = self.complex_idx(m, -1, m_complex, l_harmonic)
            for col in plus.tolist():
                to_m[row, col] = 1.0
                row += 1
            m_size[m] = len(plus)
            for col in minus.tolist():
                to_m[row, col] = 1.0
                row += 1
        self.register_buffer("l_harmonic", l_harmonic)
        self.register_buffer("m_harmonic", m_harmonic)
        self.register_buffer("m_complex", m_complex)
        self.register_buffer("res_size", torch.tensor(sizes, dtype=torch.long))
        self.register_buffer("to_m", to_m.detach())
        self.register_buffer("m_size", m_size)
        self.lmax_cache = self.mmax_cache = None
        self.mask_indices_cache = None
        self.rotate_inv_rescale_cache = None

    def complex_idx(self, m, lmax, m_complex, l_harmonic):
        if lmax == -1:
            lmax = max(self.lmax_list)
        idx = torch.arange(len(l_harmonic), device=l_harmonic.device)
        ok = l_harmonic.le(lmax)
        plus = idx[ok & m_complex.eq(m)]
        minus = idx[ok & m_complex.eq(-m)] if m != 0 else idx[:0]
        return plus, minus

    def coefficient_idx(self, lmax, mmax):
        if self.lmax_cache == lmax and self.mmax_cache == mmax and self.mask_indices_cache is not None:
            return self.mask_indices_cache
        keep = self.l_harmonic.le(lmax) & self.m_harmonic.le(mmax)
        self.device = keep.device
        self.mask_indices_cache = torch.arange(len(keep), device=keep.device)[keep]
        self.lmax_cache, self.mmax_cache = lmax, mmax
        self.rotate_inv_rescale_cache = None
        return self.mask_indices_cache

    def get_rotate_inv_rescale(self, lmax, mmax):
        if (self.lmax_cache == lmax and self.mmax_cache == mmax and self.rotate_inv_rescale_cache is not None):
            return self.rotate_inv_rescale_cache
        mask = self.coefficient_idx(lmax, mmax)
        K = (lmax + 1) ** 2
        scale = torch.ones(1, K, K, device=mask.device)
        for l in range(mmax + 1, lmax + 1):
            lo, hi = l * l, (l + 1) ** 2
            scale[:, lo:hi, lo:hi] = math.sqrt((2 * l + 1) / (2 * mmax + 1))
        self.rotate_inv_rescale_cache = scale[:, :, mask]
        return self.rotate_inv_rescale_cache

    def __repr__(self):
        return f"{self.__class__.__name__}(lmax_list={self.lmax_list}, mmax_list={self.mmax_list})"


class SO3_Embedding:
    """Plain container [length, sum (lmax+1)^2, channels] with public mutable fields, as the
    reference models use it (`.embedding`, `.lmax_list`, `.mmax_list`, ...)."""

    def __init__(self, length, lmax_list, num_channels, device, dtype):
        self.num_channels = num_channels
        self.device = device
        self.dtype = dtype
        self.num_resolutions = len(lmax_list)
        self.num_coefficients = sum(int((l + 1) ** 2) for l in lmax_list)
        self.set_embedding(torch.zeros(length, self.num_coefficients, num_channels, device=device, dtype=dtype))
        self.set_lmax_mmax(lmax_list, lmax_list.copy())

    def _like(self, num_channels=None):
        return SO3_Embedding(0, self.lmax_list.copy(), num_channels or self.num_channels, self.device, self.dtype)

    def clone(self):
        out = self._like()
        out.set_embedding(self.embedding.clone())
        return out

    def set_embedding(self, embedding):
        self.length = len(embedding)
        self.embedding = embedding

    def set_lmax_mmax(self, lmax_list, mmax_list):
        self.lmax_list = lmax_list
        self.mmax_list = mmax_list

    def _expand_edge(self, edge_index):
        self.set_embedding(self.embedding[edge_index])

    def expand_edge(self, edge_index):
        out = self._like()
        out.set_embedding(self.embedding[edge_index])
        return out

    def _reduce_edge(self, edge_index, num_nodes):
        """Sum edge rows onto nodes (reference: index_add_, so3.py:304-318) -- here through the
        deterministic dst-sorted order."""
        order = torch.sort(edge_index, stable=True)[1]
        out = torch.zeros(num_nodes, self.embedding.shape[1], self.embedding.shape[2],
                          device=self.embedding.device, dtype=self.embedding.dtype)
        out.index_add_(0, edge_index[order], self.embedding[order])
        self.set_embedding(out)

    def _m_primary(self, mapping):
        self.embedding = torch.einsum("nac,ba->nbc", self.embedding, mapping.to_m)

    def _l_primary(self, mapping):
        self.embedding = torch.einsum("nac,ab->nbc", self.embedding, mapping.to_m)

    def _rotate(self, SO3_rotation, lmax_list, mmax_list):
        parts, off = [], 0
        for i in range(self.num_resolutions):
            n = int((self.lmax_list[i] + 1) ** 2)
            parts.append(SO3_rotation[i].rotate(self.embedding[:, off:off + n], lmax_list[i], mmax_list[i]))
            off += n
        self.embedding = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        self.set_lmax_mmax(lmax_list.copy(), mmax_list.copy())

    def _rotate_inv(self, SO3_rotation, mappingReduced):
        parts, off = [], 0
        for i in range(self.num_resolutions):
            n = int(mappingReduced.res_size[i])
            parts.append(SO3_rotation[i].rotate_inv(self.embedding[:, off:off + n], self.lmax_list[i], self.mmax_list[i]))
            off += n
        self.embedding = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        self.set_lmax_mmax(self.lmax_list, [int(l) for l in self.lmax_list])

    def _grid_act(self, SO3_grid, act, mappingReduced):
        off = 0
        pieces = []
        for i in range(self.num_resolutions):
            n = int(mappingReduced.res_size[i])
            grid = SO3_grid[self.lmax_list[i]][self.mmax_list[i]]
            x = self.embedding[:, off:off + n]
            g = act(torch.einsum("bai,zic->zbac", grid.get_to_grid_mat(None), x))
            pieces.append(torch.einsum("bai,zbac->zic", grid.get_from_grid_mat(None), g))
            off += n
        self.embedding = pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=1)

    def to_grid(self, SO3_grid, lmax=-1):
        if lmax == -1:
            lmax = max(self.lmax_list)
        grid = SO3_grid[lmax][lmax]
        outs, off = [], 0
        for i in range(self.num_resolutions):
            n = int((self.lmax_list[i] + 1) ** 2)
            mat = grid.get_to_grid_mat(None)[:, :, grid.mapping.coefficient_idx(self.lmax_list[i], self.lmax_list[i])]
            outs.append(torch.einsum("bai,zic->zbac", mat, self.embedding[:, off:off + n]))
            off += n
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=3)

    def _from_grid(self, x_grid, SO3_grid, lmax=-1):
        if lmax == -1:
            lmax = max(self.lmax_list)
        grid = SO3_grid[lmax][lmax]
        outs, ch = [], 0
        for i in range(self.num_resolutions):
            mat = grid.get_from_grid_mat(None)[:, :, grid.mapping.coefficient_idx(self.lmax_list[i], self.lmax_list[i])]
            part = x_grid if self.num_resolutions == 1 else x_grid[:, :, :, ch:ch + self.num_channels]
            outs.append(torch.einsum("bai,zbac->zic", mat, part))
            ch += self.num_channels
        self.embedding = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


class SO3_Rotation(nn.Module):
    """Wigner-D matrices of the edge frames (reference so3.py:490-545)."""

    def __init__(self, lmax):
        super().__init__()
        self.lmax = lmax
        self.mapping = CoefficientMappingModule([self.lmax], [self.lmax])
        self.wigner_packed = None
        self._dense = None

    def set_wigner(self, rot_mat3x3):
        """Per-forward state, exactly as in the reference (equiformerv2_qm9.py:576-577)."""
        self.device, self.dtype = rot_mat3x3.device, rot_mat3x3.dtype
        self.wigner_packed = ops.wigner_from_rot(rot_mat3x3, self.lmax)
        self._dense = None

    def _dense_pair(self):
        if self._dense is None:
            w = ops.wigner_to_dense(self.wigner_packed, self.lmax)
            self._dense = (w, w.transpose(1, 2).contiguous())
        return self._dense

    @property
    def wigner(self):
        return self._dense_pair()[0]

    @property
    def wigner_inv(self):
        return self._dense_pair()[1]

    def rotate(self, embedding, out_lmax, out_mmax):
        mask = self.mapping.coefficient_idx(out_lmax, out_mmax)
        return torch.bmm(self.wigner[:, mask, :], embedding)

    def rotate_inv(self, embedding, in_lmax, in_mmax):
        mask = self.mapping.coefficient_idx(in_lmax, in_mmax)
        scale = self.mapping.get_rotate_inv_rescale(in_lmax, in_mmax)
        return torch.bmm(self.wigner_inv[:, :, mask] * scale, embedding)


class SO3_Grid(nn.Module):
    """S2 grid projection matrices (reference so3.py:552-646); buffers `to_grid_mat`,
    `from_grid_mat` [res_beta, res_alpha, Kr] plus `mapping.*`."""

    def __init__(self, lmax, mmax, normalization="integral", resolution=None):
        super().__init__()
        self.lmax = lmax
        self.mmax = mmax
        self.lat_resolution = 2 * (lmax + 1)
        self.long_resolution = 2 * (mmax + 1) + 1 if lmax == mmax else 2 * mmax + 1
        if resolution is not None:
            self.lat_resolution = self.long_resolution = resolution
        self.mapping = CoefficientMappingModule([lmax], [lmax])
        if normalization != "component":
            raise NotImplementedError("SO3_Grid: only normalization='component' is used by the models")
        tg, fg = _so3_math.s2_grid_matrices(lmax, mmax, self.lat_resolution, self.long_resolution)
        self.register_buffer("to_grid_mat", torch.from_numpy(tg.copy()))
        self.register_buffer("from_grid_mat", torch.from_numpy(fg.copy()))
        self._padded = {}

    def get_to_grid_mat(self, device):
        return self.to_grid_mat

    def get_from_grid_mat(self, device):
        return self.from_grid_mat

    def to_grid(self, embedding, lmax, mmax):
        mat = self.to_grid_mat[:, :, self.mapping.coefficient_idx(lmax, mmax)]
        return torch.einsum("bai,zic->zbac", mat, embedding)

    def from_grid(self, grid, lmax, mmax):
        mat = self.from_grid_mat[:, :, self.mapping.coefficient_idx(lmax, mmax)]
        return torch.einsum("bai,zbac->zic", mat, grid)

    def kernel_mats(self, order):
        """Padded [G, KP] copies of the buffers for the fused kernel, columns in l-primary ('l')
        or m-primary ('m') reduced order.  Rebuilt if the buffers were reloaded / moved."""
        key = (self.to_grid_mat.device, self.to_grid_mat._version, self.to_grid_mat.data_ptr())
        hit = self._padded.get(order)       # one entry PER ORDER: a grid with lmax == mmax serves both the attention
        if hit is None or hit[0] != key:    # ('m') and the FFN ('l') -- a single entry would be rebuilt on every call
            hit = self._padded[order] = (key, ops.GridMats.from_buffers(
                self.to_grid_mat, self.from_grid_mat, self.lmax, self.mmax, order))
        return hit[1]


class SO3_LinearV2(nn.Module):
    """Per-degree linear map, bias on l = 0 only (reference so3.py:698-743)."""

    def __init__(self, in_features, out_features, lmax, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.lmax = lmax
        bound = 1 / math.sqrt(in_features)
        self.weight = nn.Parameter(torch.empty(lmax + 1, out_features, in_features).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(out_features))
        self.register_buffer("expand_index", torch.tensor(
            [l for l in range(lmax + 1) for _ in range(2 * l + 1)], dtype=torch.long))

    def forward(self, input_embedding):
        out = ops.so3_linear(input_embedding.embedding, self.weight, self.bias)
        res = SO3_Embedding(0, input_embedding.lmax_list.copy(), self.out_features,
                            device=input_embedding.device, dtype=input_embedding.dtype)
        res.set_embedding(out)
        res.set_lmax_mmax(input_embedding.lmax_list.copy(), input_embedding.lmax_list.copy())
        return res

    def __repr__(self):
        return f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, lmax={self.lmax})"


class SO3_Linear(nn.Module):
    """Never instantiated by any reference model; kept as an import target."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("SO3_Linear is dead code in the reference; use SO3_LinearV2")
