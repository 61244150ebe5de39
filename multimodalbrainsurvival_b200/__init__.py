"""B200-native (sm_100a) implementation of the Cox-loss survival hot path of
gevaertlab/MultiModalBrainSurvival.  See DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401
from .cox import CoxLoss, cox_loss, risk_order  # noqa: F401
from . import dist, pipeline  # noqa: F401
from .aggregate import aggregate_case_features, get_survival_CI, segmented_mean, survival_grouping  # noqa: F401

__all__ = ["CoxLoss", "cox_loss", "risk_order", "aggregate_case_features", "get_survival_CI",
           "segmented_mean", "survival_grouping"]
