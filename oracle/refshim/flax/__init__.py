"""flax stand-in (TEST INFRASTRUCTURE, see oracle/refshim/README.md)."""
from . import core, jax_utils, linen, training  # noqa: F401

__version__ = "0.3.4-refshim"
