import sys, ctypes
sys.path.insert(0, ".")
import torch
from tests.configs import product_model
from cluster_generator_b200 import _native as N
pm = product_model("vikhlinin")
v = pm.dm_virial
for _ in range(2):
    v._generate_df()
torch.cuda.synchronize()
L = ctypes.CDLL(N.LIB_PATH)
buf = (ctypes.c_longlong * 8)()
L.cgb_test_debug_cycles(buf)
print("fpback cycles apply/factor", buf[0], buf[2])
