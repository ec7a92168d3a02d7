import sys; sys.path.insert(0,'.')
import numpy as np, warnings; warnings.filterwarnings("ignore")
from oracle import denovo3d_oracle as O
from helicon_b200 import pipeline as P, solver_linear_regression as S
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import CandidateSpec, MAX_EQUATIONS, positive_rule
from tests.helpers import csr_equal
d=np.load("tests/golden/task_c2_thresh.npz")
apix,twist,rise,csym,pc,tf=[float(v) for v in d["args"]]
data=d["image"].copy(); ny,nx=data.shape
nr=min(ny//2-1,int(np.ceil(ny*apix/2/apix)+1))
data-=np.median(data[(ny//2-nr,ny//2+nr),:])
data=P.threshold_data(data,thresh_fraction=tf); data/=np.max(data)
D2,D3,L2,L3=[int(v) for v in d["geom"]]
so=160
prob=Problem(data,1.0,D2,L2,D3,0.0,D3//2-1)
n3=L3*prob.ndisk; target=min(MAX_EQUATIONS,int(max(D2*L2,n3)*so))
b=Batch(prob,L3,[CandidateSpec(twist,rise/apix,int(csym),target,target,False)])
print("flags tie", b.tie.sum(), b.plan.cand_tie_z, "K/MC", b.plan.MC, "views", len(b.plan.views), "pairs", len(b.plan.pairs))
A,bb,pid=b.data_csr(0)
Ar,br,pr=O.build_A_data_matrix(data,1.0,twist,rise/apix,int(csym),0,0,0,D2,L2,D3,0,L3,target,"nn")
print("data rows", A.shape, Ar.shape, csr_equal(A,Ar), np.array_equal(bb,br))
As,bs=b.sym_csr(0)
Asr,bsr=O.build_A_helical_sym_matrix(L3,D3,D3,twist,rise/apix,int(csym),0.0,D3//2-1,target,"nn")
print("sym rows", As.shape, Asr.shape, csr_equal(As,Asr))
res=b.solve(clip_pred=1)
print(res)
(rec,_,_),score=O.lsq_reconstruct(data,1.0,twist,rise/apix,int(csym),0,0,0,tf,int(pc),0,D2,D3,L2,L3,so,"nn")
print("oracle score", score, "gpu", res[0]["score"])
x=b.x(0)
print("rel x", np.linalg.norm(b.rec3d(0)-rec)/np.linalg.norm(rec))
res2=b.solve(clip_pred=0); print("noclip gpu score", res2[0]["score"])
import scipy.sparse as sp
from scipy.sparse.linalg import lsmr
AA=sp.vstack((Ar,Asr)).tocsr(); rhs=np.concatenate((br,bsr))
r=lsmr(AA,rhs,atol=1e-4,btol=1e-4,maxiter=1000); print("scipy itn",r[2], "istop", r[1])
xr=r[0].astype(np.float32); pred=Ar.dot(xr); print("cos noclip", O.cosine_similarity(pred,br), "clip", O.cosine_similarity(np.clip(pred,0,None),br))
