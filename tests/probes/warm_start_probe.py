"""Multiplier / mu warm start (NON-REFERENCE mode: the scripts never pass lam_x0 / lam_g0) measured on the CPU
oracle: (a) re-solve at the same p from (x*, lam*), (b) the closed loop with shifted primal AND dual warm starts.
Result recorded in DESIGN.md section 7: (a) 19 -> 6..9 iterations, (b) no reduction (21.6 -> 21.5 mean, 18.3 -> 17..19 median),
whatever mu_init and the pushes are, so the kernel keeps IPOPT's cold multiplier start.
    python tests/probes/warm_start_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, oracle
import b200nmpc, bench
def shift_duals(sc, lam_x, lam_g, ok):
    B=lam_x.shape[0]; N=sc.N; R=5+sc.n_obs
    lx=lam_x.reshape(B,N,6); lx=np.concatenate([lx[:,1:],lx[:,-1:]],1).reshape(B,-1).copy()
    lg=lam_g.reshape(B,N+1,R); lg=np.concatenate([lg[:,1:],lg[:,-1:]],1).reshape(B,-1).copy()
    lx[~ok,0]=np.nan
    return lx,lg
for name,mu in [("t_trajectory",1e-4),("t_trajectory",1e-5),("t_trajectory",1e-6),("race_track_2",1e-5),("nmpc_tt",1e-3),("10_obstacles",1e-5)]:
    sc=b200nmpc.SCENARIOS[name]
    spec=oracle.make_spec(sc.T,sc.N,sc.n_obs,sc.w1,sc.w2,sc.vfov,sc.hfov); obs=sc.obstacle_table()
    lbx,ubx,lbg,ubg=sc.bounds(); B=128
    p0,vw=b200nmpc.random_instances(sc,B,seed=5)
    res={}
    for mode in ["cold","warm"]:
        p=p0.copy(); u=np.zeros((B,6*sc.N)); lx=lg=None; its=[]; oks=[]; errs=[]
        oracle.set_option("ws_mu_init",mu)
        for k in range(25):
            r=oracle.solve(spec,obs,p,u,lbx,ubx,lbg,ubg,lam_x0=lx if mode=="warm" else None,lam_g0=lg if mode=="warm" else None)
            ok=r['status']==0
            if mode=="warm": lx,lg=shift_duals(sc,r['lam_x'],r['lam_g'],ok)
            u=bench.host_shift(sc.T,p,r['x'],vw)
            its.append(r['iters'].mean()); oks.append(ok.mean())
        res[mode]=(np.mean(its[3:]),np.mean(oks[3:]),p.copy())
        oracle.clear_options()
    d=np.abs(res['cold'][2]-res['warm'][2]).max(1)
    print(name,mu,'cold its %.1f ok %.3f | warm its %.1f ok %.3f | state diff after 25 steps median %.2e p90 %.2e'%(res['cold'][0],res['cold'][1],res['warm'][0],res['warm'][1],np.median(d),np.quantile(d,.9)))
