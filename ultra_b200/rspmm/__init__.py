"""Same import surface as /root/reference/ultra/rspmm/__init__.py:1 (`from .rspmm import *`)."""
from .rspmm import *  # noqa: F401,F403
from .rspmm import (generalized_rspmm, RSPMMAddMulFunction, RSPMMMinMulFunction, RSPMMMaxMulFunction,  # noqa: F401
                    RSPMMAddAddFunction, RSPMMMinAddFunction, RSPMMMaxAddFunction)
from .pna import generalized_rspmm_pna, RSPMMPnaFunction  # noqa: F401  (one-pass PNA aggregation)
