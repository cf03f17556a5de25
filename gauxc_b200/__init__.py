"""gauxc_b200: B200-native (sm_100a) EXC/VXC quadrature hot path behind GauXC's API.

The product is gauxc_b200/libgauxc_b200.so (C ABI in include/gauxc_b200.h); this package is
the thin host-side mirror of the reference's object model over that ABI.
"""
from . import capi, systems  # noqa: F401
from .capi import (BasisSet, Functional, GauXCError, LoadBalancerFactory, MolGrid, Molecule,  # noqa: F401
                   MolecularWeightsFactory, RuntimeEnvironment, XCIntegratorFactory)
