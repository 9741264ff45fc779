"""The swap the reference needs: `from fortran_modules import particle` resolves to the GPU
module with the f2py attribute path and signatures (python_scripts/halo_gas.py:6,182,208)."""
import inspect
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_attribute_path_and_signatures():
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fortran_modules import particle\n"
        "import inspect\n"
        "f = particle.particle.brute_force_binding_energy\n"
        "g = particle.particle.serial_brute_force_binding_energy\n"
        "pf = [p.name for p in inspect.signature(f).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]\n"
        "pg = [p.name for p in inspect.signature(g).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]\n"
        "assert pf == ['ncores','ntotal','total_mass','total_x','total_y','total_z','ntest','test_x','test_y','test_z'], pf\n"
        "assert pg == pf[1:], pg\n"
        "ps = [p.name for p in inspect.signature(particle.particle.halo_shape).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]\n"
        "assert ps == ['ncore','npart','x','y','z','mass'], ps\n"          # particle_subroutines.f90:160
        "pq = [p.name for p in inspect.signature(particle.particle.sigma_projections).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]\n"
        "assert pq == ['ncore','npart','grid','n_cell','part_list','st_x','st_y','st_z','st_vx','st_vy','st_vz','st_mass',"
        "'cx','cy','cz','R05x','R05y','R05z','ll'], pq\n"                  # particle_subroutines.f90:217-219"
        "print('ok')\n" % (os.path.join(ROOT, "dropin"), ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "ok"


def test_host_mirror_signatures_match_reference_names():
    from pyhalma_b200 import halo_gas, halo_properties
    rps = [p.name for p in inspect.signature(halo_gas.RPS).parameters.values()
           if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert rps == ["gas_x", "gas_y", "gas_z", "gas_vx", "gas_vy", "gas_vz", "gas_mass", "gas_temp", "dm_x", "dm_y",
                   "dm_z", "dm_mass", "st_x", "st_y", "st_z", "st_mass", "vx", "vy", "vz", "BRUTE_FORCE_LIM",
                   "mass_dm_part", "num_dm_species"]                       # halo_gas.py:285-287
    mb = [p.name for p in inspect.signature(halo_gas.most_bound_particle).parameters.values()
          if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert mb == ["gas_x", "gas_y", "gas_z", "gas_mass", "dm_x", "dm_y", "dm_z", "dm_mass", "st_x", "st_y", "st_z",
                  "st_mass", "st_oripa", "BRUTE_FORCE_LIM", "mass_dm_part"]   # halo_gas.py:498-499
    ev = [p.name for p in inspect.signature(halo_properties.escape_velocity_unbinding_fortran).parameters.values()
          if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert ev == ["rete", "L", "ncoarse", "grid_data", "gas_data", "masclet_dm_data", "cx", "cy", "cz", "vx", "vy",
                  "vz", "Rmax", "part_list", "st_x", "st_y", "st_z", "st_vx", "st_vy", "st_vz", "st_mass",
                  "factor_v", "rho_B"]                                      # halo_properties.py:282-285
