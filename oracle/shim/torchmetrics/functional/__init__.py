import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_oracle = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _oracle not in sys.path:
    sys.path.insert(0, _oracle)
from torchmetrics_port import (  # noqa: E402,F401
    structural_similarity_index_measure,
    peak_signal_noise_ratio,
    mean_squared_error,
)
