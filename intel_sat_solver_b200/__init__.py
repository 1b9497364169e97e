"""B200-native batched Boolean-constraint-propagation engine behind the Topor API surface."""
from .topor import BatchedTopor, BatchResult, BbcpError, TToporLitVal, TToporReturnVal  # noqa: F401
