"""CPU probe (oracle): what becomes of the solves that enter IPOPT's restoration phase / watchdog along a closed loop.
DESIGN.md section 5 quotes its output: of ~1000 solves that entered the restoration phase (3 scenarios x 2048 instances x 8
closed-loop steps) two converged afterwards -- the machinery decides WHICH failure exit an infeasible or ill-conditioned NLP takes
and which iterate is handed back, it does not rescue solves.   python tests/probes/resto_outcomes.py"""
import sys, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import b200nmpc, oracle
from oracle import nlp_ref
for name in ('race_track_2','nmpc_tt','t_trajectory'):
    sc=b200nmpc.SCENARIOS[name]
    sp=oracle.make_spec(sc.T,sc.N,sc.n_obs,sc.w1,sc.w2,sc.vfov,sc.hfov)
    obs=sc.obstacle_table(); lbx,ubx,lbg,ubg=sc.bounds()
    B=2048
    p,vw=b200nmpc.random_instances(sc,B,seed=5)
    u0=np.zeros((B,sc.n_w))
    for step in range(8):
        r=oracle.solve(sp,obs,p,u0,lbx,ubx,lbg,ubg); st=r['stats']
        resto=st[:,2]>0
        hist=np.bincount(r['status'][resto],minlength=7)
        print(name,'step',step,'conv %.3f'%(r['status']==0).mean(),'resto instances',int(resto.sum()),'their status hist',hist.tolist(),
              'wd-only conv', int(((st[:,4]>0)&~resto&(r['status']==0)).sum()),'of',int(((st[:,4]>0)&~resto).sum()))
        # closed loop shift
        for b in range(B):
            x0n,un,xsn=nlp_ref.shift_timestep(sc.T,p[b,:8],r['x'][b].reshape(sc.N,6).T,p[b,8:],vw[b])
            p[b,:8]=x0n; p[b,8:]=xsn; u0[b]=un.T.reshape(-1)
