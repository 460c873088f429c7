"""TEST INFRASTRUCTURE ONLY -- import shim so the UNMODIFIED reference sources under
/root/reference can be executed in the build container (see oracle/ref_loader.py).
Re-exports the restatement in oracle/; nothing here is product code."""
from sklearn.gaussian_process.kernels import *  # noqa: F401,F403
from sklearn.gaussian_process.kernels import (  # noqa: F401
    RBF, ConstantKernel, Exponentiation, Matern, Product, Sum, WhiteKernel)
