"""Test configuration.

`-m "not gpu"` (CPU container): oracle vs golden vectors, host logic, C-ABI export check, and the
kernel SOURCE executed by the CPU emulator in tests/emu (TEST INFRASTRUCTURE: the product library is
built by nvcc only and the package refuses CPU tensors; the emulator is bound here by monkey-patching
the ctypes loader, the product knows nothing about it).
`-m gpu` (B200 box): the same parity tests through the real libeqv2_b200.so on cuda:0.
"""
import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

PKG = "equivarianttransformermpnn4quantumcomputations_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


class Backend:
    def __init__(self, name):
        self.name = name
        self.device = torch.device("cuda:0" if name == "cuda" else "cpu")

    def to(self, t):
        if isinstance(t, dict):
            return {k: self.to(v) for k, v in t.items()}
        if isinstance(t, (list, tuple)):
            return type(t)(self.to(v) for v in t)
        return t.to(self.device) if torch.is_tensor(t) else t


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """'emu': CPU tensors + kernel source under the emulator; 'cuda': the real library on cuda:0."""
    import importlib
    _lib = importlib.import_module(PKG + "._lib")
    ops = importlib.import_module(PKG + ".ops")
    if request.param == "emu":
        sys.path.insert(0, os.path.join(REPO, "tests", "emu"))
        import build_emu
        path = build_emu.build()
        monkeypatch.setitem(_lib._state, "lib", _lib._bind(path))
        monkeypatch.setattr(_lib, "check_device", lambda *t: None)
        monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
        ops.set_gemm_mode("fp32")           # the tcgen05 engine is inline PTX: GPU only
        # the software-pipelined node kernels are bit-identical to their plain twins (tests/test_rotate_pipeline.py, which
        # switches them on); 105 KB stages per emulated CTA only make the model-level emulator tests slower
        monkeypatch.setenv("EQV2_NODE_PIPE", "0")
    else:
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        monkeypatch.setitem(_lib._state, "lib", None)      # force the real library
        _lib.lib()
        ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)         # the product default (f16x3)
    ops._plan_cache.clear()
    for lay in ops.CoeffLayout._cache.values():
        lay._dev.clear()
    yield Backend(request.param)
    ops._plan_cache.clear()
    ops.set_gemm_mode(ops.DEFAULT_GEMM_MODE)


def golden(name):
    return torch.load(os.path.join(REPO, "tests", "golden", name), weights_only=False)
