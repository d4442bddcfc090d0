"""kB as the reference defines it (/root/reference/openmmtools/constants.py:7)."""
from . import unit

kB = unit.BOLTZMANN_CONSTANT_kB * unit.AVOGADRO_CONSTANT_NA   # 8.31446261815324e-3 kJ/mol/K
KB_MD = 8.31446261815324e-3
