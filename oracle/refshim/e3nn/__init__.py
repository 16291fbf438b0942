"""ORACLE shim for `e3nn` (test infrastructure). See oracle/sh_basis.py for conventions."""
from . import o3  # noqa: F401
