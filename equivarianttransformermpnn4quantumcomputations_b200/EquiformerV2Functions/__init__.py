"""Drop-in replacement for the reference package `models/EquiformerV2Functions`
(same module names, class names, constructor / forward signatures and state_dict keys;
SURVEY §8b).  The tensor work is done by hand-written sm_100a kernels through the C ABI of
include/eqv2_b200.h; there is no PyTorch/CPU fallback for the hot-path operators."""
