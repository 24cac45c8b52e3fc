"""Exception types of the routed modules (ultralytics/utils/errors.py:6-11,45-46,65-76).

When this package runs inside the reference tree the reference's own classes are re-exported, so `except MoERouterError`
written against `ultralytics.utils.errors` keeps catching what the B200 modules raise; standalone, equivalent classes with the
same names, constructor and message layout are defined here."""
from __future__ import annotations

try:  # drop-in: share class identity with the host framework when it is importable
    from ultralytics.utils.errors import MoERouterError, ShapeMismatchError, YOLOMasterError  # type: ignore
except Exception:  # noqa: BLE001 - any import problem means "standalone"

    class YOLOMasterError(Exception):
        """Base class of the errors raised by YOLO-Master modules."""

    class MoERouterError(YOLOMasterError):
        """A routed module received an invalid input or configuration."""

    class ShapeMismatchError(YOLOMasterError):
        """A routed tensor violates an expected shape contract."""

        def __init__(self, expected, actual, context: str = ""):
            self.expected, self.actual, self.context = expected, actual, context
            message = f"Shape mismatch: expected {expected}, got {actual}"
            if context:
                message += f" [{context}]"
            super().__init__(message)

__all__ = ["YOLOMasterError", "MoERouterError", "ShapeMismatchError"]
