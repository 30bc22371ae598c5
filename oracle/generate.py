"""Inputs are produced by the driver-side generators (hpddm_b200/examples/generate.py, the
mirror of the reference's examples/generate.cpp); re-exported here for the oracle's tests."""
from hpddm_b200.examples.generate import *  # noqa: F401,F403
from hpddm_b200.examples.generate import generate2d, generate3d, generate_world, split_grid_3d  # noqa: F401
