"""Seeded synthetic structures of each BASELINE config's shape.  Host-side numpy; the
batch dicts follow the reference collate schema {atomic_numbers, pos, batch, natoms[, cell, pbc]}
(data_loader_oc20v2.py:172-200, data_loader_qm9_v4.py:242-256)."""
import numpy as np
import torch


def oc20_slab(rng, n_extra=16):
    """4-layer 4x4 fcc(100)-like metal slab (64 atoms) + `n_extra` light atoms above it, cell ~10.2 x 10.2 x 35 A."""
    a = 2.55
    metal = int(rng.choice([13, 22, 26, 27, 28, 29, 45, 46, 47, 78, 79]))
    pts = []
    for k in range(4):
        shift = 0.5 * a * (k % 2)
        for i in range(4):
            for j in range(4):
                pts.append([i * a + shift, j * a + shift, 8.0 + k * 1.80])
    pos = np.array(pts) + rng.normal(0, 0.05, (64, 3))
    Z = [metal] * 64
    extra = []
    while len(extra) < n_extra:
        p = np.array([rng.uniform(0, 4 * a), rng.uniform(0, 4 * a), rng.uniform(8.0 + 3 * 1.8 + 1.2, 8.0 + 3 * 1.8 + 4.5)])
        allp = np.concatenate([pos, np.array(extra).reshape(-1, 3)])
        d = allp - p
        d[:, 0] -= np.round(d[:, 0] / (4 * a)) * 4 * a
        d[:, 1] -= np.round(d[:, 1] / (4 * a)) * 4 * a
        if np.sqrt((d ** 2).sum(1)).min() > 0.9:
            extra.append(p)
    pos = np.concatenate([pos, np.array(extra)])
    Z += [int(z) for z in rng.choice([1, 6, 7, 8], n_extra)]
    cell = np.diag([4 * a, 4 * a, 35.0]) + rng.normal(0, 0.02, (3, 3))
    return np.array(Z), pos, cell


def oc20_batch(num_structures, seed):
    rng = np.random.default_rng(seed)
    Z, pos, cell, batch, natoms = [], [], [], [], []
    for g in range(num_structures):
        z, p, c = oc20_slab(rng)
        Z.append(z); pos.append(p); cell.append(c)
        batch.append(np.full(len(z), g)); natoms.append(len(z))
    data = dict(atomic_numbers=torch.from_numpy(np.concatenate(Z)).long(),
                pos=torch.from_numpy(np.concatenate(pos)).float(),
                batch=torch.from_numpy(np.concatenate(batch)).long(),
                natoms=torch.tensor(natoms, dtype=torch.long),
                cell=torch.from_numpy(np.stack(cell)).float())
    n = data["pos"].shape[0]
    g2 = torch.Generator().manual_seed(seed + 1)
    data["energy"] = torch.randn(num_structures, generator=g2)
    data["forces"] = 0.1 * torch.randn(n, 3, generator=g2)
    return data


def qm9_batch(num_molecules, seed, nmin=9, nmax=29):
    """Molecules of U{nmin..nmax} atoms, uniform in a cube of side (9 A^3 n)^(1/3), Z in {H,C,N,O,F},
    min pair distance > 0.7 A."""
    rng = np.random.default_rng(seed)
    Z, pos, batch, natoms = [], [], [], []
    for g in range(num_molecules):
        n = int(rng.integers(nmin, nmax + 1))
        side = (9.0 * n) ** (1.0 / 3.0)
        pts = []
        while len(pts) < n:
            p = rng.uniform(0, side, 3)
            if not pts or np.sqrt(((np.array(pts) - p) ** 2).sum(1)).min() > 0.7:
                pts.append(p)
        Z.append(rng.choice([1, 6, 7, 8, 9], n)); pos.append(np.array(pts))
        batch.append(np.full(n, g)); natoms.append(n)
    return dict(atomic_numbers=torch.from_numpy(np.concatenate(Z)).long(),
                pos=torch.from_numpy(np.concatenate(pos)).float(),
                batch=torch.from_numpy(np.concatenate(batch)).long(),
                natoms=torch.tensor(natoms, dtype=torch.long))


def matpes_batch(num_cells, seed, n_atoms=30, vol_per_atom=15.0):
    """Bulk cells of `n_atoms` atoms at 15 A^3/atom, lattice = L*I + 5 % Gaussian shear, Z ~ U{1..89}
    (cfg 3/4); structures with a pair closer than 0.7 A (any of the 27 images) are resampled."""
    rng = np.random.default_rng(seed)
    Z, pos, cell, batch, natoms = [], [], [], [], []
    side = (vol_per_atom * n_atoms) ** (1.0 / 3.0)
    shifts = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], dtype=float)
    for g in range(num_cells):
        lat = side * np.eye(3) + 0.05 * side * rng.normal(size=(3, 3))
        pts = []
        while len(pts) < n_atoms:
            p = rng.uniform(0, 1, 3) @ lat
            ok = True
            if pts:
                d = (np.array(pts)[:, None, :] + (shifts @ lat)[None, :, :]) - p
                ok = np.sqrt((d ** 2).sum(-1)).min() > 0.7
            if ok:
                pts.append(p)
        Z.append(rng.integers(1, 90, n_atoms)); pos.append(np.array(pts)); cell.append(lat)
        batch.append(np.full(n_atoms, g)); natoms.append(n_atoms)
    data = dict(atomic_numbers=torch.from_numpy(np.concatenate(Z)).long(),
                pos=torch.from_numpy(np.concatenate(pos)).float(),
                batch=torch.from_numpy(np.concatenate(batch)).long(),
                natoms=torch.tensor(natoms, dtype=torch.long),
                cell=torch.from_numpy(np.stack(cell)).float())
    g2 = torch.Generator().manual_seed(seed + 1)
    data["energy"] = torch.randn(num_cells, 1, generator=g2)
    data["forces"] = 0.1 * torch.randn(num_cells * n_atoms, 3, generator=g2)
    return data
