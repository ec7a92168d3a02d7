import sys, os; sys.path.insert(0,'.')
import numpy as np, warnings; warnings.filterwarnings("ignore")
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import CandidateSpec, MAX_EQUATIONS
d=np.load("tests/golden/solve_nn_unb_64.npz")
apix, twist, rise, csym, pc, so, L3 = d["args"]; img=d["image"]; N=img.shape[0]; L3=int(L3)
prob=Problem(img,1.0,N,N,N,0.0,N//2-1)
n3=L3*prob.ndisk; target=min(MAX_EQUATIONS,int(max(N*N,n3)*so))
b=Batch(prob,L3,[CandidateSpec(float(twist),float(rise/apix),1,target,target,False)])
rng=np.random.default_rng(0)
x=rng.standard_normal(b.n).astype(np.float32)
y=b.apply_forward(0,x)
np.save(f"gpurun_out/y_{os.environ.get('HB2_NO_DEDUPE','0')}.npy",y)
g=b.apply_adjoint(0,y); np.save(f"gpurun_out/g_{os.environ.get('HB2_NO_DEDUPE','0')}.npy",g)
res=b.solve(fixed_iters=1, check_every=1)
print(os.environ.get('HB2_NO_DEDUPE'), "fwd |y|", np.linalg.norm(y), "res", res["score"], res["normr"], res["normA"], res["normar"])
v=b.plan.views
c0=b.plan.cands[0]
print([(i,int(q["dup_of"])) for i,q in enumerate(v) if q["dup_of"]>=0])
