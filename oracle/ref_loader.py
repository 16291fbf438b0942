"""ORACLE tooling: import the UNMODIFIED reference (/root/reference/models) in the build
container.  /root/reference does not exist on the GPU box, so nothing that runs there may
call this; it is used only by oracle/make_golden.py and the CPU-side pinning tests (which
skip when the reference tree is absent)."""
import os
import sys

import torch

REF_ROOT = "/root/reference/models"
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
            if not getattr(mod, "__file__", "").startswith(REF_ROOT):
                del sys.modules[name]
