"""clu stand-in (only parameter_overview is imported by the reference)."""
from . import parameter_overview  # noqa: F401
