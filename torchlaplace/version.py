from neurallaplace_b200.version import __version__  # noqa: F401
