"""Drop-in replacement package for pyHALMA's `fortran_modules` directory."""
