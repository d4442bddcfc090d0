#!/usr/bin/env python
"""Golden Sobol' points from the reference's own generator (runs standalone: numpy only).

LennardJonesFluid initial positions are ``float32(i4_sobol_generate(3, N, 1)) * box_edge``
(/root/reference/openmmtools/testsystems.py:277-284).  Run in the build container only.
"""
import os, sys, importlib.util
import numpy as np
spec = importlib.util.spec_from_file_location('ref_sobol', '/root/reference/openmmtools/sobol.py')
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
out = {}
for n in (1, 8, 100, 512, 1000):
    out[f'sobol3_n{n}_skip1'] = np.array(m.i4_sobol_generate(3, n, 1), dtype=np.float64)   # shape (3, n)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sobol_golden.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst))
print(out['sobol3_n8_skip1'].T)
