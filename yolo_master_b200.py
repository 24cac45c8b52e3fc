"""Import shim: the package directory is `yolo-master_b200/` (not a valid identifier), so
`import yolo_master_b200` resolves here and is redirected to that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "yolo-master_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
