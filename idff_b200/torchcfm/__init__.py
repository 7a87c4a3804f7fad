"""Drop-in surface of the reference's `torchcfm` package for the IDFF hot path (torchcfm/__init__.py:1-2)."""
from .conditional_flow_matching import *  # noqa: F401,F403
from .version import __version__  # noqa: F401
