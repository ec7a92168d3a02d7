import sys; sys.path.insert(0,'.')
import numpy as np, warnings; warnings.filterwarnings("ignore")
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import CandidateSpec, MAX_EQUATIONS
np.set_printoptions(linewidth=200, precision=6, suppress=True)
rng = np.random.default_rng(11)
N, L3 = 40, 4
yy, xx = np.mgrid[0:N, 0:N]
img = np.zeros((N, N), np.float32)
for _ in range(25):
    cy, cx = rng.uniform(12, 28), rng.uniform(0, N)
    img += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 5.0).astype(np.float32)
prob = Problem(img, 1.0, N, N, N, 0.0, N//2-1)
target = min(MAX_EQUATIONS, int(max(N*N, L3*prob.ndisk)*4))
b = Batch(prob, L3, [CandidateSpec(-2.37, 1.31, 1, target, target, True)])
res = b.solve()
print(res)
tr, nit, status = b.trf_trace(0)
print("nit", nit, "status", status); print(tr)
