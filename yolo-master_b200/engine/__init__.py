"""Predictor-side steps either side of the forward pass (mirror of `ultralytics.engine`, detection inference subset)."""
from .predictor import DetectionPredictor
from .results import Boxes, Results

__all__ = ["DetectionPredictor", "Results", "Boxes"]
