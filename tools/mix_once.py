"""One swap-all call on the K=256 LJ energy matrix (for ncu captures of the walker).  usage: mix_once.py [nswap]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmmtools_b200._engine import Engine
K = 256
u = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'u_lj_256.npy'))
e = Engine(0, K, K, 0)
e.set_energies(u); e.set_replica_states(np.arange(K)); e.mix_seed(1234, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else K ** 3 // 8
e.mix_swap_all(n)
print(e.mix_stats())
