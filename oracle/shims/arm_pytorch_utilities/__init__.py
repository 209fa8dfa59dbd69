"""Shim: see oracle/shims/README.md and oracle/tp_arm_utils.py."""
from oracle.tp_arm_utils import tensor_utils, rand  # noqa: F401
