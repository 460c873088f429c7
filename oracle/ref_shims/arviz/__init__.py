"""TEST INFRASTRUCTURE ONLY -- import shim so the UNMODIFIED reference sources under
/root/reference can be executed in the build container (see oracle/ref_loader.py).
Re-exports the restatement in oracle/; nothing here is product code."""


def hdi(*args, **kwargs):
    raise NotImplementedError("arviz.hdi is off the hot path (bask/optimizer.py:685)")
