"""Drop-in for the reference's ``src/models/fortitran.py`` (:10, :253)."""
from adafortitran_b200.estimators import BaseFortiTranEstimator, FortiTranEstimator

__all__ = ["BaseFortiTranEstimator", "FortiTranEstimator"]
