"""``torchlaplace.version`` of the drop-in package (reference: ``/root/reference/torchlaplace/version.py``)."""
from neurallaplace_b200.version import __version__  # noqa: F401
