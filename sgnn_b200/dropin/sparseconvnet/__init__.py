"""Drop-in `sparseconvnet` module: put  <repo>/sgnn_b200/dropin  (and <repo>) on sys.path and the
reference's `import sparseconvnet as scn` (torch/model.py:7) binds to the B200 engine."""
from sgnn_b200.scn import *          # noqa: F401,F403
from sgnn_b200.scn import __all__    # noqa: F401
