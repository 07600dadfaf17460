import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import default_params
from vicon2gt_b200 import sim, api
from vicon2gt_b200._cdefs import TrialInfo
from oracle import binding as ob
np.set_printoptions(linewidth=220, precision=10)
ds=sim.make_config("C2",seed=7,trajectory="synth", duration_s=25.0,vicon_sigmas=(1e-3,1e-3,1e-3,1e-2,1e-2,1e-2))
p=default_params()
for faithful in (1,0):
    o=ob.OracleSolver.from_dataset(p,ds); o.set_option("faithful_linerr",faithful); o.build_and_solve()
    tr_o=o.lm_trace()
    print("oracle faithful",faithful,"iters",o.info(TrialInfo).iterations,"tries",len(tr_o), "final", repr(o.info(TrialInfo).final_error))
    if faithful==1: tr_o1=tr_o
b=api.BatchSolver(1); b.set_trial_dataset(0,p,ds); b.build_and_solve()
tr_g=b.lm_trace(0)
print("gpu iters",b.info(0).iterations,"tries",len(tr_g),"final",repr(b.info(0).final_error))
n=max(len(tr_o),len(tr_g))
for i in range(n):
    a=tr_o[i] if i<len(tr_o) else None; g=tr_g[i] if i<len(tr_g) else None
    f=lambda r: "%.0e %.12e %.12e %.6e %d"%(r[0],r[1],r[2],r[3],r[4]) if r is not None else "-"
    print(i, f(a), "|", f(g))
