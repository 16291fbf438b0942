"""ORACLE stand-in for fairchem.core.graph.compute.generate_graph (call site
equiformerv2_oc20.py:223-234).  fairchem is un-vendored and un-pinned; this brute-force
restatement follows its documented semantics: all periodic images within `cutoff`,
edge_index[0] = neighbour j, edge_index[1] = centre i, vec = pos[j] - pos[i] + offset,
per-centre truncation on SQUARED distances: keep d2 <= d2_sorted[max_neighbors] + 0.01 when
enforce_max_neighbors_strictly=False (fairchem-core `get_max_neighbors_mask`), self images need d2 > 1e-4.
PARITY UNPINNED: restated from the published fairchem-core source, not checked against a run of it."""
import torch

from oracle.eqv2_oracle import radius_graph_pbc_fairchem


def generate_graph(data, cutoff, max_neighbors, enforce_max_neighbors_strictly=False,
                   radius_pbc_version=1, pbc=None):
    dev = data.pos.device          # the brute-force restatement runs on the host whatever device the model lives on
    ei, dist, vec = radius_graph_pbc_fairchem(
        data.pos.detach().cpu(), data.cell.detach().cpu(), data.batch.cpu(), data.natoms.cpu(), cutoff, max_neighbors,
        enforce_max_neighbors_strictly)
    return {"edge_index": ei.to(dev), "edge_distance": dist.to(dev), "edge_distance_vec": vec.to(dev)}
