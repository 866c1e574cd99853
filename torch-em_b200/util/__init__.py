from .prediction import Blocking, predict_with_halo, standardize

__all__ = ["Blocking", "predict_with_halo", "standardize"]
