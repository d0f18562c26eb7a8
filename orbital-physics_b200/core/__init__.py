"""B200-native drop-in for the `core` package of trevormcguire/orbital-physics.

Same import surface (core.engine, core.physics, core.body, core.units,
core.constants, core.datasets, core.examples, core.plot); the per-timestep hot
path -- all-pairs gravity + leapfrog step -- runs as sm_100a CUDA kernels behind
the C ABI in include/orbital_b200.h (see core/_native.py).  No CPU fallback.
"""
