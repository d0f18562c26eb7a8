"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's hot path (core/physics.py:125-159,
core/engine.py:65-97, core/physics.py:391-422,510-535 of
trevormcguire/orbital-physics).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package, and
only as the checker / the timed CPU arm.  Parity status: PINNED against outputs
of the unmodified reference (tests/golden/, tests/test_oracle.py).
"""
from .c_oracle import COracle, load as load_c_oracle  # noqa: F401
from . import ref_numpy  # noqa: F401
