"""ORACLE shim for torch_geometric (only utils.softmax is used: transformer_block.py:315)."""
from . import utils  # noqa: F401
