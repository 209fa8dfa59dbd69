"""Do-nothing matplotlib so that the reference's visualization.py imports (plotting is out of scope)."""
from . import pyplot, colors  # noqa: F401
