import sys
import numpy as np
from parity_common import build_pair
ctx, orc = build_pair(dim=2, s=1, ref=3, n=2, ell=1, stabilize=len(sys.argv) > 1)
try:
    X, Minv, G = ctx.debug_stages(0, want_G=False)
    res = orc.compute_patch(0)
    print(np.round(X[:3], 6))
    print("Minv", Minv)
except Exception as e:
    print("FAIL", e)
