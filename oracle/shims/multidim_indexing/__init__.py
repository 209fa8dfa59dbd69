"""Shim: see oracle/shims/README.md and oracle/tp_multidim_indexing.py."""
from . import torch_view  # noqa: F401
