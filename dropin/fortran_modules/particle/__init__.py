"""`from fortran_modules import particle` -> `particle.particle.<routine>` (python_scripts/halo_gas.py:6,182;
halo_properties.py:13,805,861).

brute_force_binding_energy / serial_brute_force_binding_energy run on the B200 through
libhalma_unbind.so; halo_shape / sigma_projections are forwarded to the original f2py build
when it is kept next to this package as `fortran_modules/_particle_f2py*.so`."""
from pyhalma_b200.particle import particle  # noqa: F401
