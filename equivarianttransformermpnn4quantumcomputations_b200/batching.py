"""Input side of the hot path: pinned-memory collate of the reference's batch dict and padding of a
prepared batch to (atoms, edges) BUCKETS, so that a CUDA-graph capture keyed by the padded sizes serves every batch
that falls into the bucket (real data has a new (atoms, edges) pair almost every step).

  collate(structures, pin=True) : list of per-structure dicts -> batch dict in the reference schema
                                  (data_loader_matpes.py:290-314 `collate_matpes`: atomic_numbers, pos, cell, energy,
                                  forces, natoms, batch, pbc [, stress, magmom]) assembled directly in page-locked buffers,
                                  so `tensor.to(device, non_blocking=True)` is a true asynchronous copy.
  pad_to_bucket(batch, n_mult, e_mult) : append ONE ghost structure whose atoms and edges fill the batch up to the next
                                  multiples of (n_mult, e_mult).

Why a ghost structure is exact: structures are independent graphs (block-diagonal edge_index; every reduction of the
path is per destination node or per graph).  The ghost's atoms only have edges among themselves, so no real
atom receives a message from them; the loss masks the ghost's energy / forces (`structure_mask`, `atom_mask`), so no
gradient flows into its activations and its rows add exact zeros to every parameter gradient.  Real outputs and
parameter gradients are those of the unpadded batch (tests/test_batching.py: <= 1e-6 relative, the difference being
summation order and the f16x3 engine's per-tensor operand scale).
"""
import torch

PER_ATOM = ("atomic_numbers", "pos", "batch", "forces", "magmom", "tags", "fixed")
PER_STRUCTURE = ("natoms", "cell", "energy", "stress", "pbc", "targets")
PER_EDGE = ("edge_distance", "edge_distance_vec", "edge_frames")

GHOST_SPACING = 1.7          # Angstrom between consecutive ghost atoms (every ghost edge joins neighbours on a line)
GHOST_ORIGIN = 1.0e3


def collate(structures, pin=True):
    """structures: list of dicts with 'atomic_numbers' [n], 'pos' [n,3], 'cell' [3,3] (optional), 'energy' [1] / scalar,
    'forces' [n,3] (optional), 'stress' [6] (optional), 'magmom' [n] (optional).  Returns the reference batch dict; all
    tensors live in pinned host memory when `pin` (one allocation per key, filled slice by slice)."""
    B = len(structures)
    counts = [int(s["atomic_numbers"].shape[0]) for s in structures]
    N = sum(counts)
    use_pin = bool(pin) and torch.cuda.is_available()

    def buf(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=use_pin)

    out = {"atomic_numbers": buf((N,), torch.long), "pos": buf((N, 3), torch.float32), "batch": buf((N,), torch.long),
           "natoms": buf((B,), torch.long)}
    first = structures[0]
    if "forces" in first:
        out["forces"] = buf((N, 3), torch.float32)
    if "magmom" in first:
        out["magmom"] = buf((N,), torch.float32)
    if "cell" in first:
        out["cell"] = buf((B, 3, 3), torch.float32)
        out["pbc"] = buf((B, 3), torch.bool)
        out["pbc"].fill_(True)
    if "energy" in first:
        out["energy"] = buf((B, 1), torch.float32)
    if "stress" in first:
        out["stress"] = buf((B, 6), torch.float32)
    off = 0
    for i, (s, n) in enumerate(zip(structures, counts)):
        sl = slice(off, off + n)
        out["atomic_numbers"][sl] = s["atomic_numbers"]
        out["pos"][sl] = s["pos"]
        out["batch"][sl] = i
        out["natoms"][i] = n
        if "forces" in out:
            out["forces"][sl] = s["forces"]
        if "magmom" in out:
            out["magmom"][sl] = s["magmom"] if s.get("magmom") is not None else 0.0
        if "cell" in out:
            out["cell"][i] = s["cell"].reshape(3, 3)
        if "energy" in out:
            out["energy"][i] = torch.as_tensor(s["energy"], dtype=torch.float32).reshape(-1)[0]
        if "stress" in out:
            out["stress"][i] = s["stress"].reshape(6)
        off += n
    return out


def bucket_sizes(N, E, n_mult, e_mult):
    """Padded (atoms, edges): the ghost structure needs >= 2 atoms (its edges join two distinct atoms)."""
    n_pad = -(-(N + 2) // n_mult) * n_mult
    e_pad = -(-E // e_mult) * e_mult
    return n_pad, e_pad


def pad_to_bucket(batch, n_mult=64, e_mult=512):
    """batch: a PREPARED batch on the device (the model's `prepare()` output merged in: `edge_index` [2,E] sorted by
    destination, optionally `edge_distance`, `edge_distance_vec`, `edge_frames`).  Returns a new dict padded with one ghost
    structure to bucket_sizes(N, E, ...), plus `atom_mask` [N_pad] / `structure_mask` [B+1] (1 = real) for the loss.
    Per-atom / per-structure / per-edge tensors are recognised by key (PER_ATOM, PER_STRUCTURE, PER_EDGE)."""
    from . import ops
    pos = batch["pos"]
    dev = pos.device
    N, B = int(pos.shape[0]), int(batch["natoms"].shape[0])
    ei = batch["edge_index"]
    E = int(ei.shape[1])
    n_pad, e_pad = bucket_sizes(N, E, n_mult, e_mult)
    ng, eg = n_pad - N, e_pad - E
    out = dict(batch)
    ghost_pos = torch.zeros(ng, 3, dtype=pos.dtype, device=dev)
    ghost_pos[:, 0] = GHOST_ORIGIN + GHOST_SPACING * torch.arange(ng, device=dev, dtype=pos.dtype)
    ghost_pos[:, 1:] = GHOST_ORIGIN
    fill = {"atomic_numbers": 1, "batch": B}
    for k in PER_ATOM:
        v = batch.get(k)
        if not torch.is_tensor(v):
            continue
        if k == "pos":
            out[k] = torch.cat([v, ghost_pos])
        else:
            out[k] = torch.cat([v, torch.full((ng,) + tuple(v.shape[1:]), fill.get(k, 0), dtype=v.dtype, device=dev)])
    for k in PER_STRUCTURE:
        v = batch.get(k)
        if not torch.is_tensor(v):
            continue
        if k == "natoms":
            extra = torch.tensor([ng], dtype=v.dtype, device=dev)
        elif k == "cell":
            extra = (1.0e2 * torch.eye(3, dtype=v.dtype, device=dev)).unsqueeze(0)
        elif k == "pbc":
            extra = torch.ones((1,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        else:
            extra = torch.zeros((1,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        out[k] = torch.cat([v, extra])
    # ghost edges: atom N + j -> atom N + j + 1 along the line, emitted in destination order (the real edges end before
    # the first ghost destination, so the padded list stays destination-sorted)
    if eg > 0:
        j = (torch.arange(eg, device=dev) * (ng - 1)) // eg              # non-decreasing in [0, ng - 2]
        g_src, g_dst = N + j, N + j + 1
        out["edge_index"] = torch.cat([ei, torch.stack([g_src, g_dst]).to(ei.dtype)], dim=1)
        vec = torch.zeros(eg, 3, dtype=pos.dtype, device=dev)
        vec[:, 0] = GHOST_SPACING
        if torch.is_tensor(batch.get("edge_distance")):
            out["edge_distance"] = torch.cat([batch["edge_distance"], torch.full((eg,), GHOST_SPACING, dtype=pos.dtype,
                                                                                 device=dev)])
        if torch.is_tensor(batch.get("edge_distance_vec")):
            out["edge_distance_vec"] = torch.cat([batch["edge_distance_vec"], vec])
        if torch.is_tensor(batch.get("edge_frames")):
            out["edge_frames"] = torch.cat([batch["edge_frames"], ops.edge_frames(vec, None)])
    out["atom_mask"] = torch.cat([torch.ones(N, dtype=pos.dtype, device=dev), torch.zeros(ng, dtype=pos.dtype, device=dev)])
    out["structure_mask"] = torch.cat([torch.ones(B, dtype=pos.dtype, device=dev), torch.zeros(1, dtype=pos.dtype, device=dev)])
    return out


def masked_mean(x, mask):
    """Mean of x over the real rows (mask [rows], broadcast over trailing dims) -- equals x[:n_real].mean()."""
    m = mask.view((-1,) + (1,) * (x.dim() - 1))
    per_row = x[0].numel() if x.dim() > 1 else 1
    return (x * m).sum() / (mask.sum() * per_row)
