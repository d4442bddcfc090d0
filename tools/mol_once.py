"""One k_propagate_mol launch on AlanineDipeptideVacuum (for ncu).  usage: mol_once.py [K [n_steps]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmmtools_b200 import testsystems, unit, _lib
from openmmtools_b200._engine import Engine
K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
a = testsystems.AlanineDipeptideVacuum()
x = np.ascontiguousarray(a.positions.value_in_unit(unit.nanometer), np.float64)
e = Engine(_lib.RX_SYSTEM_MOLECULE, K, K, 22)
e.set_molecule(a.system)
e.set_states(np.linspace(300.0, 600.0, K), np.ones(K))
e.set_integrator(0.002, 5.0, n_steps, 'V R O R V')
e.set_positions(np.stack([x] * K)); e.randomize_velocities(3)
e.propagate(1, 1)
e.phase_times(reset=True)
e.propagate(1, 2)
print(e.phase_times())
