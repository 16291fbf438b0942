"""Drop-in replacements for the reference's `models/NewFunctions/<fork>/` packages that BASELINE configs 4-5
use.  They reuse the kernels behind `EquiformerV2Functions`."""
