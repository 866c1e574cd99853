from .prediction import Blocking, predict_with_halo, predict_with_halo_pipelined, standardize

__all__ = ["Blocking", "predict_with_halo", "predict_with_halo_pipelined", "standardize"]
