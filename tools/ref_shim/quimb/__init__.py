"""Stand-in for quimb (see ../README.md): only `quimb.tensor` is populated."""
from . import tensor  # noqa: F401

__all__ = []  # `from quimb import *` in the reference brings in nothing it uses on the executed paths
