"""The dependency-free HDF5 reader lives in the package (icp_rust_b200/hdf5_min.py); kept importable
from here for tests/golden/make_fixtures.py."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from icp_rust_b200.hdf5_min import *  # noqa: F401,F403
from icp_rust_b200.hdf5_min import MiniH5  # noqa: F401
