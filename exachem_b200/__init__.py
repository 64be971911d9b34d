"""B200-native fused CCSD(T) triples driver -- drop-in for ExaChem's exachem/cc/ccsd_t hot path.

The product is the C-ABI library ``libccsdt_b200.so`` (hand-written sm_100a CUDA + a C++ host
driver, sources in ``exachem_b200/csrc``, interface in ``include/ccsdt_b200.h``).  This package is
the thin Python mirror of the reference's operator interface used by tests and bench.py.
"""
from .driver import (CCSD_T_Fused_Driver, CcsdtError, Options, TiledSpace, lib, setup_mo_space)  # noqa: F401
