"""Which shared-memory element does the contraction MMA of the S/V kernel read for accumulator row r, position k?
One non-zero entity j0 per launch: O[r][d] / E[j0][d] = the A value the MMA saw at (r, j0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime

N, D, G = 256, 512, 256
dbgs = [int(x) for x in sys.argv[1:]] or [12, 36]
rows = [0, 1, 8, 63, 64, 95, 96, 97, 103, 104, 111, 112, 120, 127, 128, 224, 255]
for dbg in dbgs:
    print("dbg", dbg, "(8: A[r][k] = r + 1, 32: A[r][k] = k + 1); columns = rows", rows)
    for j0 in (0, 1, 7, 8, 15, 16, 31, 32, 63, 64, 100, 127, 128, 129, 200, 255):
        ent = np.zeros((N, D), np.float32)
        ent[j0] = 1.0
        ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
        ctx.set_option("sv_dbg", dbg)
        m, l, O = ctx.contract(np.zeros((G, D), np.float32), 0)
        torch.cuda.synchronize()
        O = O.cpu().numpy()
        spread = float(np.abs(O - O[:, :1]).max())
        print(f"  j0={j0:3d}: seen", np.round(O[rows, 0], 2).tolist(), "col spread", spread, "col 300:", np.round(O[[0, 96, 127], 300], 2).tolist())
        ctx.close()
