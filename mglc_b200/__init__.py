"""mglc_b200 -- B200-native hot path of cheryli/MGLC (collide+stream lattice update, boundary kernels,
3-D subdomain halo exchange) behind the C ABI of include/mglc.h.

Everything that computes lives in libmglc.so (hand-written sm_100a CUDA, built by
`__graft_entry__.build()` / `make -C mglc_b200/csrc`).  This package is the thin host-side mirror of the
reference driver's interface; it has no CPU or PyTorch fallback and raises if the library is missing.
"""
from . import _lib, formats
from ._lib import MglcError, lib
from .jacobi import Jacobi, dims_create_nd
from .lbm_aa import LidDrivenCavityAA
from .lid2d import LidDrivenCavity2D
from .particles import ParticleChannel
from .thermal2d import BuoyancyDrivenCavity2D
from .lbm import (BuoyancyDrivenCavity, Communicator, LidDrivenCavity, make_thermal_desc, Subdomain, cart_neighbors, decompose_1d, dims_create,
                  halo_plan, halo_plan_2d, make_desc)

__all__ = ["MglcError", "lib", "BuoyancyDrivenCavity", "Communicator", "LidDrivenCavity", "make_thermal_desc", "Subdomain", "cart_neighbors",
           "decompose_1d", "dims_create", "halo_plan", "halo_plan_2d", "make_desc", "Jacobi", "dims_create_nd", "ParticleChannel", "LidDrivenCavity2D", "LidDrivenCavityAA", "BuoyancyDrivenCavity2D", "_lib", "formats"]
