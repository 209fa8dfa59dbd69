"""Shim: see oracle/shims/README.md and oracle/tp_open3d.py."""
from oracle.tp_open3d import io, geometry, t, utility  # noqa: F401
import types as _types
visualization = _types.SimpleNamespace(draw_geometries=lambda *a, **k: None)
