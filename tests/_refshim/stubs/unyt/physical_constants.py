"""TESTS ONLY -- unyt.physical_constants stand-in (CODATA 2014 mks values, as unyt's)."""
from . import unyt_quantity

G = unyt_quantity(6.67408e-11, "m**3/kg/s**2")
gravitational_constant = G
mp = unyt_quantity(1.672621898e-27, "kg")
mass_proton = mp
me = unyt_quantity(9.10938356e-31, "kg")
amu = unyt_quantity(1.660539040e-27, "kg")
mh = unyt_quantity(1.007947 * 1.660539040e-27, "kg")
mass_hydrogen = mh
kboltz = unyt_quantity(1.38064852e-23, "J/K")
boltzmann_constant = kboltz
clight = unyt_quantity(2.99792458e8, "m/s")
