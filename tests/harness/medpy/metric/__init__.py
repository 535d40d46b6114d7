from . import binary  # noqa: F401
