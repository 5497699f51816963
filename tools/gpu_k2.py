"""K2 (BCP-only replay) of the config-2 batch, for an ncu capture of replay_smem (profiling aid).
usage: gpu_k2.py [n_inst]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200pkg
pkg = _b200pkg.load()
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = pkg.Engine(0)
eng.generate_ksat(0x5EED000000000000, n_inst, 100, 426, 3)
print(eng.bcp_replay(reps=3))
