"""Predictor-side steps either side of the forward pass (mirror of `ultralytics.engine`, inference subset)."""
from .predictor import DetectionPredictor, OBBPredictor, PosePredictor, SegmentationPredictor
from .results import OBB, Boxes, Keypoints, Masks, Results

__all__ = ["DetectionPredictor", "SegmentationPredictor", "OBBPredictor", "PosePredictor", "Results", "Boxes", "Masks", "OBB", "Keypoints"]
