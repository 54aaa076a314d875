"""nanograd_b200 — a B200-native (sm_100a) device backend behind nanograd's own Tensor /
Function / Module / Optimizer API.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"

from nanograd_b200.device import Device  # noqa: F401
from nanograd_b200.tensor import Tensor  # noqa: F401
