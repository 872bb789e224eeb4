"""Diagnostic: where (variable, node, element parity) does the tuned headline kernel differ from the generic path?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import trixi_b200 as T
from elixirs import ELIXIRS

semi = ELIXIRS["tree_3d_euler_ec"].semi()
u = T.compute_coefficients(0.0, semi)
gpu = semi.backend()
res = {}
for path in (1, 0):
    gpu.set_option(gpu.OPT_KERNEL_PATH, path)
    du = np.full_like(u, np.nan)
    T.rhs_hyperbolic(du, u, semi, 0.1)
    res[path] = du.copy()
    gpu.upload(0, u)
    gpu.calc_volume_integral() if hasattr(gpu, "calc_volume_integral") else None
err = np.abs(res[0] - res[1])
scale = np.abs(res[1]).max()
print("max rel err", err.max() / scale, "nan count", np.isnan(res[0]).sum())
print("by variable", err.max(axis=(1, 2, 3, 4)) / scale)
print("by i", err.max(axis=(0, 2, 3, 4)) / scale)
print("by j", err.max(axis=(0, 1, 3, 4)) / scale)
print("by k", err.max(axis=(0, 1, 2, 4)) / scale)
ee = err.max(axis=(0, 1, 2, 3)) / scale
print("even elements", ee[0::2].max(), "odd elements", ee[1::2].max(), "bad elements", int((ee > 1e-10).sum()), "of", ee.size)
bad = np.argwhere(err / scale > 1e-10)
print("first bad entries (v,i,j,k,e):", bad[:12].tolist())
