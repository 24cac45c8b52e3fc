"""Device-side data transforms on the inference path (mirror of `ultralytics.data.augment`, predictor subset)."""
from .augment import LetterBox

__all__ = ["LetterBox"]
