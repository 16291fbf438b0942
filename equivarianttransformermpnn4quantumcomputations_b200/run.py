"""Launch a reference script on the B200 path:  python -m equivarianttransformermpnn4quantumcomputations_b200.run /path/to/models/train_x.py

Reference scripts import the hot path by bare name (`from EquiformerV2Functions.so3 import ...`,
models/equiformerv2_qm9.py:58-75) and their own directory precedes PYTHONPATH, so the drop-in is
installed by pre-populating `sys.modules` before the script runs."""
import importlib
import runpy
import sys

_SUBMODULES = ("so3", "so2_ops", "transformer_block", "activation", "layer_norm", "radial_function", "input_block",
               "edge_rot_mat", "drop", "module_list", "wigner")


def install_alias():
    pkg = importlib.import_module(__package__ + ".EquiformerV2Functions")
    sys.modules["EquiformerV2Functions"] = pkg
    for sub in _SUBMODULES:
        sys.modules["EquiformerV2Functions." + sub] = importlib.import_module(pkg.__name__ + "." + sub)
    # forks used by the GATA models: `from NewFunctions.Gotennet_morethaninspired.transformer_block import ...`
    # (models/equiformerv2_MatPES_GATAV2.py:49)
    nf = importlib.import_module(__package__ + ".NewFunctions")
    sys.modules["NewFunctions"] = nf
    for fork in ("Gotennet_morethaninspired", "Gotennets_GATA_phi_refined_every_layer", "GATA_and_all2all"):
        sys.modules["NewFunctions." + fork] = importlib.import_module(nf.__name__ + "." + fork)
        for sub in ("transformer_block", "activation"):
            try:
                sys.modules[f"NewFunctions.{fork}.{sub}"] = importlib.import_module(f"{nf.__name__}.{fork}.{sub}")
            except ModuleNotFoundError:
                pass        # GATA_and_all2all ships only `activation` (its transformer_block is unused)
    return pkg


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m equivarianttransformermpnn4quantumcomputations_b200.run <reference_script.py> [args...]")
    install_alias()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
