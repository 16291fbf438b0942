"""ORACLE tooling (TEST INFRASTRUCTURE): import the UNMODIFIED reference.

In the build container that is `/root/reference/models`; on the GPU box (no /root/reference) it is the
byte-identical copy `oracle/_ref/models` that `oracle/make_ref.py` made here and gpurun shipped
(git-ignored, never part of the history).  Used by oracle/make_golden.py, the parity tests (checker) and
bench.py's CPU legs (baseline) -- never by the product path."""
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/models" if os.path.isdir("/root/reference/models") else os.path.join(_HERE, "_ref", "models")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(REF_ROOT)


def install(lmax_jd=8):
    """Make `import EquiformerV2Functions...` / `import equiformerv2_qm9` resolve to the
    reference sources, with third-party stand-ins and a regenerated Jd."""
    if not available():
        raise RuntimeError("reference tree not present")
    for p in (_REPO, _SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import sh_basis
    jd_list = [torch.tensor(J, dtype=torch.float64) for J in sh_basis.make_jd(lmax_jd)]
    if not getattr(torch.load, "_eqv2_patched", False):
        real_load = torch.load

        def patched(f, *a, **k):
            if isinstance(f, str) and f.endswith("Jd.pt"):
                return jd_list
            return real_load(f, *a, **k)

        patched._eqv2_patched = True
        torch.load = patched
    # a product package of the same name must not shadow the reference here
    for name in list(sys.modules):
        if name == "EquiformerV2Functions" or name.startswith("EquiformerV2Functions."):
            mod = sys.modules[name]
            if not (getattr(mod, "__file__", None) or "").startswith(REF_ROOT):
                del sys.modules[name]
