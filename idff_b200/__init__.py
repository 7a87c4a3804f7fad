"""idff_b200 -- B200-native (sm_100a) implementation of the IDFF hot path.

    from idff_b200.torchcfm.conditional_flow_matching import ExactOptimalTransportConditionalFlowMatcher
    from idff_b200.torchcfm.models.models import MLP, MLP_Embedding
    from idff_b200.fields import CustomNetModel, CustomNetModelOrder3, CFMModel, SDE_cal
    from idff_b200.sampler import odeint, sdeint
    from idff_b200.mmd import compute_mmd_rbf

All compute goes through libidff_b200.so (include/idff_b200.h); there is no CPU fallback."""
from . import _lib  # noqa: F401
from ._lib import IdffError, build, launch_count  # noqa: F401

__version__ = "0.1.0"
