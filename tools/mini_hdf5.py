"""Kept for the fixture tools: the reader now lives in the package (jqmc_b200/hdf5_lite.py)."""

from jqmc_b200.hdf5_lite import MiniHDF5  # noqa: F401
