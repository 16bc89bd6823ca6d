import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel
H="HomotopySolver{SimpleSolver}"; HC="HomotopySolver{CachingSolver{SimpleSolver}}"
m=ex.birdie(vol=0.8)
B,N=64,4410
rng=np.random.default_rng(0xACE5EED)
u=np.asfortranarray(np.clip(0.2*rng.standard_normal((1,N,B)),-1,1))
yexact=OracleModel(m,B,solver=H,tol=1e-13).run(u,threads=0)
yref=OracleModel(m,B,solver=H).run(u,threads=0)
for tol in (1e-13,0.0):
    for k in ("auto","generic"):
        r=BatchRunner(m,B,solver=H,tol=tol,kernel=k); y=r.run(u); st=r.stats(); r.close()
        peak=np.abs(yexact).max(); sc=np.maximum(np.abs(yexact),1e-3*peak)
        print('tol',tol,k,'gpu-exact abs',np.abs(y-yexact).max(),'rel',(np.abs(y-yexact)/sc).max(),'gpu-ref abs',np.abs(y-yref).max(),'E_ref',np.abs(yref-yexact).max(), 'iters', st['newton_iters'], st['not_converged'], st['homotopy_solves'])
