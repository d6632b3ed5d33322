"""Import alias: `import golfer_b200` -> the package directory
`computer-vision-system-for-analyzing-golfer-action_b200/` (its name is not a
valid Python identifier, so it is loaded through importlib).  Use attribute
access (`golfer_b200.config`, `golfer_b200.segment`), not `import golfer_b200.x`.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("computer-vision-system-for-analyzing-golfer-action_b200")
sys.modules[__name__] = _pkg
