import sys, numpy as np
sys.path.insert(0, '.')
from tests.helpers import load, csr_from, csr_equal, cases
from helicon_b200 import solver_linear_regression as S
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import CandidateSpec, MAX_EQUATIONS
from oracle import denovo3d_oracle as O
from scipy.sparse import vstack
import warnings; warnings.filterwarnings("ignore")
for name in cases("data","nn"):
    d=load(name); s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl = d["args"]
    prob = Problem(d["image"], float(s), int(D2), int(L2), int(D2), D3i / 2, int(D3) // 2 - 1)
    batch = Batch(prob, int(L3), [CandidateSpec(twist, rise, int(csym), int(mpl), -1, False)])
    A,b,pid=batch.data_csr(0); ref=csr_from(d)
    print(name, "K?", "tie_xy", int(batch.tie.sum()), "tie_z", batch.plan.cand_tie_z[0], A.shape, ref.shape, csr_equal(A,ref), "b eq", len(b)==len(d["b"]) and np.array_equal(b,d["b"]))
    batch.close(); prob.close()
for name in cases("hsym","nn"):
    d=load(name); nz, ny, nx, twist, rise, csym, rmin, rmax, msp = d["args"]
    A,b=S.build_A_helical_sym_matrix(int(nz), int(ny), int(nx), float(twist), float(rise), int(csym), rmin, int(rmax), int(msp), "nn")
    ref=csr_from(d); print(name, A.shape, ref.shape, csr_equal(A,ref))
for name in ["solve_nn_unb_32","solve_nn_unb_48_t35","solve_nn_unb_48_c2","solve_nn_unb_64"]:
    d=load(name); apix, twist, rise, csym, pc, so, L3 = d["args"]; img=d["image"]; N=img.shape[0]
    (rec,_,_),score,info=S.lsq_reconstruct(img,1.0,float(twist),float(rise/apix),int(csym),positive_constraint=0,reconstruct_diameter_2d_pixel=N,reconstruct_length_2d_pixel=N,reconstruct_diameter_3d_pixel=N,reconstruct_length_3d_pixel=int(L3),sym_oversample=int(so),return_info=True)
    ref=d["rec3d"]; print(name,"itn",info["res"]["itn"],"istop",info["res"]["istop"],"score",float(score),float(d["score"]),"rel",np.linalg.norm(rec-ref)/np.linalg.norm(ref), info["res"], info["timing"])
    # operator check
    prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1)
    target = min(MAX_EQUATIONS, int(max(N * N, int(L3) * prob.ndisk) * int(so)))
    batch = Batch(prob, int(L3), [CandidateSpec(twist, rise/apix, int(csym), target, target, False)])
    A_d,b_d,_=O.build_A_data_matrix(img,1.0,float(twist),float(rise/apix),int(csym),0,0,0,N,N,N,0,int(L3),target,"nn")
    A_s,_=O.build_A_helical_sym_matrix(int(L3),N,N,float(twist),float(rise/apix),int(csym),0.0,N//2-1,target,"nn")
    pidx,kk,jj=batch.data_row_index(0); nd,tot=batch.rows_padded(0)
    print("  rows", len(pidx), A_d.shape, tot-nd, A_s.shape, "tie", int(batch.tie.sum()), batch.plan.cand_tie_z)
    if len(pidx)==A_d.shape[0] and tot-nd==A_s.shape[0]:
        x=np.random.default_rng(0).standard_normal(batch.n).astype(np.float32)
        y=batch.apply_forward(0,x); yr=A_d@x; ys=A_s@x
        print("  fwd data err", np.abs(y[:nd][pidx]-yr).max(), "scale", np.abs(yr).max(), "sym err", np.abs(y[nd:]-ys).max())
        ur=np.random.default_rng(1).standard_normal(len(pidx)+A_s.shape[0]).astype(np.float32)
        u=np.zeros(tot,np.float32); u[:nd][pidx]=ur[:len(pidx)]; u[nd:]=ur[len(pidx):]
        g=batch.apply_adjoint(0,u); gr=vstack((A_d,A_s)).T@ur
        print("  adj err", np.abs(g-gr).max(), "scale", np.abs(gr).max())
        gd=batch.apply_adjoint(0,np.concatenate((u[:nd],np.zeros(tot-nd,np.float32)))); print("  adj data-only err", np.abs(gd-A_d.T@ur[:len(pidx)]).max())
    batch.close(); prob.close()
