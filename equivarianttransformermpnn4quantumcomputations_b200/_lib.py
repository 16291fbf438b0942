"""ctypes binding of libeqv2_b200.so (C ABI declared in include/eqv2_b200.h).

There is NO fallback: if the CUDA library is missing the package raises on first use, and
every operator refuses tensors that do not live on a CUDA device.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libeqv2_b200.so")

P = ctypes.c_void_p
I = ctypes.c_int
L = ctypes.c_longlong
F = ctypes.c_float
D = ctypes.c_double


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("A", P), ("B", P), ("C", P), ("bias", P),
        ("M", I), ("N", I), ("K", I), ("transA", I), ("transB", I),
        ("a_rpb", L), ("a_bs", L), ("a_ld", L),
        ("b_rpb", L), ("b_bs", L), ("b_ld", L),
        ("c_rpb", L), ("c_bs", L), ("c_ld", L),
        ("accumulate", I),
    ]


class SplitDesc(ctypes.Structure):
    _fields_ = [("src", P), ("dst", P), ("absmax", P), ("rows", L), ("cols", L), ("rows_pad", L), ("cols_pad", L),
                ("absmax_given", I), ("slab_k", I)]


class Gemm16Desc(ctypes.Structure):
    _fields_ = [
        ("A", P), ("B", P), ("C", P), ("bias", P), ("a_absmax", P), ("b_absmax", P),
        ("a_ld", L), ("a_plane", L), ("b_ld", L), ("b_plane", L), ("c_ld", L), ("c_rpb", L), ("c_bs", L),
        ("M", I), ("N", I), ("K", I), ("transA", I), ("transB", I), ("accumulate", I), ("c_absmax", P),
    ]


MAX_GEMM_GROUPS = 10
MAX_SPLIT_ITEMS = 16

_PROTOS = {
    "eqv2_abi_version": [],
    "eqv2_gemm_f32": [P, I, I, P],
    "eqv2_gemm_tc": [P, I, I, I, P],
    "eqv2_split_f16": [P, I, P],
    "eqv2_gemm_f16": [P, I, I, P],
    "eqv2_gemm_f16_ex": [P, I, I, I, P],
    "eqv2_wigner_from_rot": [P, P, P, L, I, P],
    "eqv2_edge_frames": [P, P, P, L, I, P, P],
    "eqv2_gather_rotate_fwd": [P, P, P, P, P, P, P, P, L, I, I, I, I, I, P, P],
    "eqv2_gather_rotate_dx": [P, P, P, P, P, P, P, P, L, I, I, I, I, I, P],
    "eqv2_gather_rotate_drad": [P, P, P, P, P, P, L, I, I, I, I, I, P, P],
    "eqv2_rotinv_reduce_fwd": [P, P, P, P, P, P, P, L, I, I, L, I, I, I, F, P],
    "eqv2_rotinv_reduce_bwd": [P, P, P, P, P, P, P, P, L, I, I, L, I, I, I, F, P, P],
    "eqv2_planes_colsum": [P, L, L, L, L, I, I, P, P, P, P],
    "eqv2_gather_rotate_fwd_planes": [P, P, P, P, P, P, L, L, P, P, P, L, I, I, I, I, I, P],
    "eqv2_gather_rotate_drad_planes": [P, P, P, P, P, P, L, L, P, P, P, L, I, I, I, I, I, P],
    "eqv2_rotinv_reduce_bwd_planes": [P, P, P, P, P, P, L, L, P, F, P, P, L, I, I, L, I, I, I, F, P],
    "eqv2_s2act_padded_rows": [I],
    "eqv2_s2act_fwd": [P, L, P, L, P, L, P, P, L, I, I, I, I, I, P],
    "eqv2_s2act_bwd": [P, L, P, L, P, L, P, L, P, L, P, P, L, I, I, I, I, I, P],
    "eqv2_s2sep_supported": [I, I],
    "eqv2_s2sep_set_tables": [P, I, I, P],
    "eqv2_s2sep_fwd": [P, L, P, L, P, L, L, I, I, I, I, I, P, P],
    "eqv2_s2sep_bwd": [P, L, P, L, P, L, P, L, P, L, L, I, I, I, I, I, P, P],
    "eqv2_s2sep_bwd2": [P, L, P, L, P, L, P, L, P, L, P, L, P, L, P, L, L, I, I, I, I, I, P],
    "eqv2_attn_alpha_fwd": [P, L, P, P, P, P, P, P, P, L, L, I, I, F, P],
    "eqv2_attn_alpha_bwd": [P, L, P, P, P, P, P, P, P, P, P, L, P, P, P, L, L, I, I, F, P, P],
    "eqv2_attn_alpha_bwd2": [P, L, P, P, P, P, P, P, P, P, P, L, P, P, L, P, P, P, P, P, L, L, I, I, F, P],
    "eqv2_equiv_norm_fwd": [P, P, P, P, P, P, L, I, I, I, P, P, F, P, P],
    "eqv2_equiv_norm_bwd": [P, P, P, P, P, P, P, P, L, I, I, I, P, P, P],
    "eqv2_equiv_norm_bwd2": [P, P, P, P, P, P, P, P, P, L, I, I, I, P, P, P],
    "eqv2_rbf_fwd": [P, P, L, I, P, F, P],
    "eqv2_rbf_bwd": [P, P, P, L, I, P, F, P],
    "eqv2_rbf_bwd2": [P, P, P, P, P, L, I, P, F, P],
    "eqv2_edge_sh": [P, P, L, I, P],
    "eqv2_rbf_linear_fwd": [P, P, P, P, P, P, P, P, P, L, I, I, F, F, F, I, P],
    "eqv2_rbf_linear_chunk": [],
    "eqv2_rbf_linear_wgrad": [P, P, P, P, P, P, P, P, P, L, I, I, F, P],
    "eqv2_ln_silu_fwd": [P, P, P, P, L, I, F, P],
    "eqv2_ln_silu_bwd": [P, P, P, P, P, P, P, L, I, F, P],
    "eqv2_ln_silu_bwd2": [P, P, P, P, P, P, P, P, P, L, I, F, P],
    "eqv2_graph_ptr": [P, P, I, P],
    "eqv2_exclusive_scan": [P, P, I, P],
    "eqv2_radius_graph": [P, P, P, L, F, I, I, P, P, P, P, P, P, P, P],
    "eqv2_pbc_reps": [P, D, P, I, P],
    "eqv2_radius_graph_pbc": [P, P, P, P, P, L, D, I, I, I, P, P, P, P, P, P, P, P],
    "eqv2_radius_graph_pbc27": [P, P, P, P, L, F, I, I, I, P, P, P, P, P, P, P, P, P],
    "eqv2_csr_from_index": [P, L, L, P, P, P, P, P],
    "eqv2_segment_sum_fwd": [P, L, P, P, L, I, P],
    "eqv2_segment_sum_bwd": [P, P, P, L, P],
    "eqv2_so2_block_weight": [P, P, I, I, P],
    "eqv2_so2_block_weight_adj": [P, P, I, I, P],
    "eqv2_embed_rows": [P, P, P, L, I, P],
    "eqv2_seg_colsum": [P, L, P, P, L, I, I, I, P, P, P],
    "eqv2_htr_inner": [P, P, P, P, L, I, I, P],
    "eqv2_htr_grad": [P, P, P, P, L, I, I, P],
    "eqv2_gata_value_fwd": [P, P, P, P, L, I, I, I, I, P],
    "eqv2_gata_value_bwd": [P, P, P, P, P, P, L, I, I, I, I, P],
    "eqv2_gata_value_bwd2": [P, P, P, P, P, P, P, P, P, L, I, I, I, I, P],
    "eqv2_s2sep_fwd_planes": [P, L, P, L, P, L, L, P, F, P, L, I, I, I, I, I, P],
    "eqv2_drop_path_scale": [P, P, P, F, P, L, L, P],
    "eqv2_pair_scores": [P, P, P, P, P, P, L, I, I, I, F, P],
    "eqv2_pair_mix": [P, P, P, P, P, P, L, I, I, I, I, I, P],
    "eqv2_pair_softmax_fwd": [P, P, P, P, L, I, P],
    "eqv2_pair_softmax_bwd": [P, P, P, P, P, L, I, P],
    "eqv2_pair_softmax_bwd2": [P, P, P, P, P, P, P, L, I, P],
    "eqv2_opt_chunk_elems": [],
    "eqv2_grad_sqnorm": [P, P, P, I, F, P, P, P],
    "eqv2_adamw_ema_step": [P, P, P, I, P, F, F, F, I, F, P],
}
# entry points that only exist in the real (nvcc-built) library
_OPTIONAL = {"eqv2_gemm_tc", "eqv2_split_f16", "eqv2_gemm_f16", "eqv2_gemm_f16_ex", "eqv2_gather_rotate_fwd_planes",
             "eqv2_gather_rotate_drad_planes", "eqv2_rotinv_reduce_bwd_planes", "eqv2_planes_colsum",
             "eqv2_s2sep_fwd_planes"}   # inline-PTX kernels: not part of the CPU emulator build

_state = {"lib": None, "launches": 0}


class Eqv2Error(RuntimeError):
    pass


def _bind(path):
    lib = ctypes.CDLL(path)
    lib.eqv2_last_error.restype = ctypes.c_char_p
    for name, args in _PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in _OPTIONAL:
                continue
            raise
        fn.argtypes = args
        fn.restype = I
    return lib


def lib():
    if _state["lib"] is None:
        if not os.path.exists(LIB_PATH):
            raise Eqv2Error(
                f"{LIB_PATH} is missing: build it with `python -m equivarianttransformermpnn4quantumcomputations_b200.build` "
                "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this package.")
        _state["lib"] = _bind(LIB_PATH)
    return _state["lib"]


def launch_count():
    return _state["launches"]


def reset_launch_count():
    _state["launches"] = 0


def check_device(*tensors):
    """Every tensor handed to a kernel must live on a CUDA device."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise Eqv2Error("eqv2_b200 kernels need CUDA tensors; there is no CPU fallback")


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_timing = {"on": False, "records": []}


def start_kernel_timing():
    """Measurement aid (bench.py): bracket every C-ABI call with CUDA events on the launching stream
    and remember the algorithmic work the caller attached (`work=(flops, bytes)`)."""
    _timing["on"] = True
    _timing["records"] = []


def stop_kernel_timing():
    """-> {entry point: dict(ms, calls, flops, bytes)} (synchronises)."""
    _timing["on"] = False
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, work in _timing["records"]:
        r = out.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
        r["ms"] += e0.elapsed_time(e1)
        r["calls"] += 1
        if work is not None:
            r["flops"] += work[0]
            r["bytes"] += work[1]
    _timing["records"] = []
    return out


def call(name, *args, n_kernels=1, work=None):
    fn = getattr(lib(), name)
    if _timing["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*args)
    if rc != 0:
        raise Eqv2Error(f"{name} failed ({rc}): {lib().eqv2_last_error().decode()}")
    if _timing["on"]:
        e1.record()
        _timing["records"].append((name, e0, e1, work))
    _state["launches"] += n_kernels
    return rc
