"""B200-native Graph Neural Solver hot path (drop-in for ref GNS/main.py:107-202).

Public surface (mirrors the reference module):

* ``GNS`` / ``LearningBlock`` - the ``nn.Module`` with the reference constructor,
  attributes and ``state_dict`` keys, running on hand-written sm_100a kernels.
* ``get_BLG`` - the column maps (ref GNS/utils.py:4-13).
* ``TopologyPlan`` - the one-time CSR plan behind the kernels.
* ``data`` - packing / perturbation helpers (ref GNS/utils.py:17-41, GNS/augment_grids.py:25-54).
* ``train`` / ``evaluate`` - the pieces of the reference's training and evaluation scripts that touch
  the hot path (ref GNS/main.py:243-309, GNS/evaluate.py:15-18,73-148), batched.

There is no CPU implementation in this package: without the CUDA library the
module raises at call time.
"""
from .model import GNS, LearningBlock, get_BLG            # noqa: F401
from .plan import TopologyPlan                            # noqa: F401
from . import data, evaluate, parallel, train             # noqa: F401
from ._lib import load_library, library_path, build_library  # noqa: F401

__all__ = ["GNS", "LearningBlock", "get_BLG", "TopologyPlan", "data", "evaluate", "parallel", "train",
           "load_library", "library_path", "build_library"]
