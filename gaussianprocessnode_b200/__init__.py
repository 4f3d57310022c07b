"""gaussianprocessnode_b200 -- B200-native data sweep for sparse variational GP factor nodes.

Only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA + the C ABI, built into ``libsgp.so``), the ctypes
binding of that ABI (``_lib`` / ``sgp``) and the host-side mirror of the reference's node interface (``nodes``) plus the N-sharding helpers (``shard``)."""
from .sgp import SGPContext, SGPError, pinned_empty, SE, MATERN32, MATERN52, SRCUBATURE, GENUT, GAUSSHERMITE, CLOSED_FORM_SE, POINT  # noqa: F401
from . import nodes  # noqa: F401
from . import shard  # noqa: F401
from . import theta  # noqa: F401
