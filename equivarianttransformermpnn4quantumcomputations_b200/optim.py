"""Optimizer-side step of the reference training loops as ONE multi-tensor kernel pass.

Reference (train_oc20v2_parallel.py:95-126,177-186,444-462): `clip_grad_norm_(params, grad_clip)` ->
`AdamW(add_weight_decay(...)).step()` -> `ExponentialMovingAverage.update(named_parameters)`.  Here:

    opt = FusedAdamW(param_groups, lr=..., weight_decay=..., max_grad_norm=grad_clip, ema_decay=0.999)
    loss.backward(); opt.step()                      # norm + clip + AdamW + EMA: 3 kernel launches for the whole model
    ema = opt.ema(model.named_parameters())          # object with the reference EMA's .shadow / store / restore / copy_to

`FusedAdamW` is a `torch.optim.Optimizer`: parameter groups (per-group lr / weight_decay as produced by the reference's
`add_weight_decay`), LR schedulers, `zero_grad`, `state_dict` (per-parameter `step`, `exp_avg`, `exp_avg_sq` under the
names torch.optim.AdamW uses, plus `ema`) all behave as usual.  The kernels (csrc/optim.cu) read ONE device table of
tensor descriptors; it is rebuilt only when a gradient pointer or a learning rate changed.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

_TABLE_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("ema", "<u8"), ("n", "<i8"),
                         ("lr", "<f4"), ("wd", "<f4"), ("lag", "<i4"), ("pad", "<i4")])   # == eqv2_opt_tensor, 64 bytes


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0,
                 ema_decay=0.0):
        if not 0.0 <= ema_decay < 1.0:
            raise ValueError("ema_decay must be in [0, 1)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        betas_set = {tuple(g["betas"]) for g in self.param_groups}
        eps_set = {g["eps"] for g in self.param_groups}
        if len(betas_set) != 1 or len(eps_set) != 1:
            raise ValueError("FusedAdamW: betas / eps must be the same in every parameter group")
        self.max_grad_norm = float(max_grad_norm)
        self.ema_decay = float(ema_decay)
        self._steps = 0
        self._taken = {}                 # id(param) -> optimizer steps it has taken (torch: per-parameter `step`)
        self._plan = None
        self._key = None
        self.grad_norm = None            # device tensor [2]: (global gradient norm before clipping, clip coefficient)

    # -- layout -------------------------------------------------------------------------------------------
    def _params(self):
        return [(p, g) for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _build_plan(self, plist):
        dev = plist[0][0].device
        chunk = int(_lib.lib().eqv2_opt_chunk_elems())
        ct, ci = [], []
        for t, (p, _) in enumerate(plist):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.Eqv2Error("FusedAdamW: parameters must be contiguous fp32 tensors")
            n = (p.numel() + chunk - 1) // chunk
            ct += [t] * n
            ci += list(range(n))
        pin = dev.type == "cuda"
        host = torch.empty(len(plist) * _TABLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=pin)
        self._plan = dict(
            device=dev, nchunks=len(ct),
            chunk_tensor=torch.tensor(ct, dtype=torch.int32, device=dev),
            chunk_index=torch.tensor(ci, dtype=torch.int32, device=dev),
            host=host, host_np=host.numpy().view(_TABLE_DTYPE),
            table=torch.empty(len(plist) * _TABLE_DTYPE.itemsize, dtype=torch.uint8, device=dev),
            partial=torch.empty(len(ct), dtype=torch.float32, device=dev),
            params=[p for p, _ in plist])
        self.grad_norm = torch.ones(2, dtype=torch.float32, device=dev)

    def _state_of(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if self.ema_decay > 0.0:
                st["ema"] = p.detach().clone()
        return st

    # -- step ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plist = self._params()
        if not plist:
            return loss
        if self._plan is None or len(self._plan["params"]) != len(plist) or \
                any(a is not b for a, (b, _) in zip(self._plan["params"], plist)):
            self._build_plan(plist)
        plan = self._plan
        _lib.check_device(*[p for p, _ in plist])
        grads = [p.grad for p, _ in plist]
        key = (tuple(0 if g is None else g.data_ptr() for g in grads), tuple(g["lr"] for g in self.param_groups),
               tuple(g["weight_decay"] for g in self.param_groups))
        if key != self._key:
            tab = plan["host_np"]
            for i, ((p, grp), g) in enumerate(zip(plist, grads)):
                st = self._state_of(p)
                if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.is_sparse):
                    raise _lib.Eqv2Error("FusedAdamW: gradients must be dense contiguous fp32 tensors")
                tab[i] = (p.data_ptr(), 0 if g is None else g.data_ptr(), st["exp_avg"].data_ptr(),
                          st["exp_avg_sq"].data_ptr(), st["ema"].data_ptr() if "ema" in st else 0, p.numel(),
                          float(grp["lr"]), float(grp["weight_decay"]), self._steps - self._taken.get(id(p), 0), 0)
            plan["table"].copy_(plan["host"], non_blocking=True)
            self._key = key
        self._steps += 1
        for (p, _), g in zip(plist, grads):
            if g is not None:
                self._taken[id(p)] = self._taken.get(id(p), 0) + 1
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        stream = _lib.stream_ptr()
        clip = None
        n_elems = float(sum(p.numel() for p, _ in plist))
        if self.max_grad_norm > 0.0:
            _lib.call("eqv2_grad_sqnorm", plan["table"].data_ptr(), plan["chunk_tensor"].data_ptr(),
                      plan["chunk_index"].data_ptr(), plan["nchunks"], self.max_grad_norm, plan["partial"].data_ptr(),
                      self.grad_norm.data_ptr(), stream, n_kernels=2, work=(0.0, 4.0 * n_elems))
            clip = self.grad_norm.data_ptr()
        _lib.call("eqv2_adamw_ema_step", plan["table"].data_ptr(), plan["chunk_tensor"].data_ptr(),
                  plan["chunk_index"].data_ptr(), plan["nchunks"], clip, float(b1), float(b2), float(eps), self._steps,
                  self.ema_decay, stream,
                  work=(0.0, (28.0 + (8.0 if self.ema_decay > 0.0 else 0.0)) * n_elems))
        return loss

    def state_dict(self):
        for p, _ in self._params():             # per-parameter `step` entries are materialised on demand
            if p in self.state and "step" in self.state[p]:
                self.state[p]["step"].fill_(float(self._taken.get(id(p), 0)))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._taken = {id(p): int(self.state[p]["step"]) for p, _ in self._params() if p in self.state}
        self._steps = max(self._taken.values(), default=0)
        self._key = None

    # -- EMA view with the reference class's interface -------------------------------------------------------
    def ema(self, named_parameters):
        """Object with `.shadow` {name: tensor} and store / restore / copy_to, as the reference's
        ExponentialMovingAverage (train_oc20v2_parallel.py:95-126); `.update` is a no-op -- `step()` already did it."""
        if self.ema_decay <= 0.0:
            raise ValueError("this optimizer was built with ema_decay = 0")
        named = [(n, p) for n, p in named_parameters if p.requires_grad]
        return _EmaView(self, named)


class _EmaView:
    def __init__(self, opt, named):
        self._opt, self._named = opt, named
        self.decay = opt.ema_decay
        self.backup = {}

    @property
    def shadow(self):
        return {n: self._opt._state_of(p)["ema"] for n, p in self._named}

    def update(self, parameters=None):
        return None

    def store(self, parameters=None):
        self.backup = {n: p.data.clone() for n, p in self._named}

    def restore(self, parameters=None):
        for n, p in self._named:
            p.data.copy_(self.backup[n])

    def copy_to(self, parameters=None):
        sh = self.shadow
        for n, p in self._named:
            p.data.copy_(sh[n])
