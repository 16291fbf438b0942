"""Shared test helpers (TEST INFRASTRUCTURE)."""
import contextlib
import importlib

import torch

from conftest import PKG


def pkg(sub):
    return importlib.import_module(PKG + "." + sub)


@contextlib.contextmanager
def fixed_rand_like(draw):
    """Make the edge-frame helper draw of init_edge_rot_mat (`torch.rand_like`, reference
    edge_rot_mat.py:28) return the recorded values, so both paths see the same frames."""
    orig = torch.rand_like

    def fake(t, *a, **k):
        assert t.shape == draw.shape, (t.shape, draw.shape)
        return draw.to(device=t.device, dtype=t.dtype)

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = orig


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def build_oc20(hp, device):
    m = pkg("models.equiformerv2_oc20").EquiformerV2_OC20(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        norm_type=hp["norm_type"], lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"],
        edge_channels=hp["edge_ch"], alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)


def build_qm9(hp, device):
    m = pkg("models.equiformerv2_qm9").EquiformerV2_QM9(
        num_targets=hp["num_targets"], max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"],
        max_num_elements=hp["max_elements"], num_layers=hp["num_layers"], sphere_channels=hp["C"],
        attn_hidden_channels=hp["H"], num_heads=hp["heads"], attn_alpha_channels=hp["alpha_ch"],
        attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"], lmax_list=[hp["lmax"]],
        mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"], alpha_drop=0.0,
        drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)


def load_params(model, params):
    own = dict(model.named_parameters())
    assert set(own) == set(params), (sorted(set(own) ^ set(params)))
    with torch.no_grad():
        for k, v in params.items():
            assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
            own[k].copy_(v)


def build_matpes_v2(hp, device):
    m = pkg("models.equiformerv2_MatPESv2").EquiformerV2_MatPES(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"],
        alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)


def build_matpes_v1(hp, device, **flags):
    m = pkg("models.equiformerv2_MatPES").EquiformerV2_MatPES(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"],
        alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0, **flags)
    return m.to(device)


def build_gatav2(hp, device):
    m = pkg("models.equiformerv2_MatPES_GATAV2").EquiformerV2_MatPES(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"],
        alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)


def build_gatav2_phi(hp, device):
    m = pkg("models.equiformerv2_MatPES_GATAV2_phi_at_every_iteration_like_gata").EquiformerV2_MatPES(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"],
        alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)


def build_gatav2_global(hp, device):
    m = pkg("models.equiformerv2_MatPES_GATAV2_GLOBALALLATTENTION_HTR_phi_at_every_iteration_like_gata_with_DISTANCE"
            ).EquiformerV2_MatPES(
        max_neighbors=hp["max_neighbors"], max_radius=hp["cutoff"], max_num_elements=hp["max_elements"],
        num_layers=hp["num_layers"], sphere_channels=hp["C"], attn_hidden_channels=hp["H"], num_heads=hp["heads"],
        attn_alpha_channels=hp["alpha_ch"], attn_value_channels=hp["value_ch"], ffn_hidden_channels=hp["ffn_hidden"],
        lmax_list=[hp["lmax"]], mmax_list=[hp["mmax"]], grid_resolution=hp["grid_res"], edge_channels=hp["edge_ch"],
        alpha_drop=0.0, drop_path_rate=0.0, proj_drop=0.0)
    return m.to(device)
