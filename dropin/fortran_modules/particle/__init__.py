"""`from fortran_modules import particle` -> `particle.particle.<routine>` (python_scripts/halo_gas.py:6,182;
halo_properties.py:13,805,861).

All four routines of the f2py module -- brute_force_binding_energy,
serial_brute_force_binding_energy, halo_shape, sigma_projections -- run on the B200 through
libhalma_unbind.so, so the swap needs no Fortran toolchain."""
from pyhalma_b200.particle import particle  # noqa: F401
