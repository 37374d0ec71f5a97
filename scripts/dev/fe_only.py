import sys, ctypes
sys.path.insert(0, ".")
import torch
from tests.configs import product_model
from cluster_generator_b200 import _native as N
pm = product_model(sys.argv[1] if len(sys.argv) > 1 else "vikhlinin")
v = pm.dm_virial
for _ in range(2):
    pm._fe_fit_cache = None
    v._generate_df()
v._generate_df()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
N.lib().cgb_test_givens_cycles(buf)
print("givens cycles factor [prologue, run, epilogue, fpback]", list(buf[:4]), "apply", list(buf[4:8]))
