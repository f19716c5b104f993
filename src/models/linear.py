"""Drop-in for the reference's ``src/models/linear.py`` (:15-107)."""
from adafortitran_b200.estimators import LinearEstimator

__all__ = ["LinearEstimator"]
