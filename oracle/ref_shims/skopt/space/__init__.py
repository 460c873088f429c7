"""TEST INFRASTRUCTURE ONLY -- import shim so the UNMODIFIED reference sources under
/root/reference can be executed in the build container (see oracle/ref_loader.py).
Re-exports the restatement in oracle/; nothing here is product code."""
from oracle.skopt_port import Categorical, Dimension, Integer, Real, Space  # noqa: F401
