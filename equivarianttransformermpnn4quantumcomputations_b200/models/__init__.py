"""Host-side model wrappers that mirror the reference's `models/equiformerv2_*.py` (same class names,
constructor arguments, forward(data: dict) contract and state_dict keys).  They are the callers of the
hot path: graph build -> edge frames -> Wigner-D -> block loop -> readout.  Tensor work is
done by the kernels behind `EquiformerV2Functions`; the wrappers hold parameters and glue."""
