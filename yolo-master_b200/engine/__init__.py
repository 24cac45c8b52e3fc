"""Predictor-side steps either side of the forward pass (mirror of `ultralytics.engine`, inference subset)."""
from .predictor import DetectionPredictor, OBBPredictor, SegmentationPredictor
from .results import OBB, Boxes, Masks, Results

__all__ = ["DetectionPredictor", "SegmentationPredictor", "OBBPredictor", "Results", "Boxes", "Masks", "OBB"]
