import torch


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """PyG segment softmax: subtract detached per-segment max, exp, divide by (segment sum + 1e-16)."""
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    shape = [n] + list(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    smax = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    smax = smax.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - smax.gather(0, idx)).exp()
    ssum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(0, idx, out)
    return out / (ssum.gather(0, idx) + 1e-16)
