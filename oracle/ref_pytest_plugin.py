"""TEST INFRASTRUCTURE ONLY -- pytest plugin: ``-p oracle.ref_pytest_plugin`` makes
``import bask`` resolve to the unmodified reference tree (see oracle/ref_loader.py), so the
reference's own test files can be collected straight from /root/reference/tests."""
from oracle.ref_loader import load_reference

load_reference()
