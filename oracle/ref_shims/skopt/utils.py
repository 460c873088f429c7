"""TEST INFRASTRUCTURE ONLY -- import shim so the UNMODIFIED reference sources under
/root/reference can be executed in the build container (see oracle/ref_loader.py).
Re-exports the restatement in oracle/; nothing here is product code."""
from oracle.skopt_port import (  # noqa: F401
    create_result, expected_minimum, is_2Dlistlike, is_listlike, normalize_dimensions)


def dimensions_aslist(search_space):
    raise NotImplementedError


def point_asdict(search_space, point_as_list):
    raise NotImplementedError
