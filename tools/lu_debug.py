import os, sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0,'/root/repo')
import wae_b200 as W
from wae_b200.nlevp import DeviceMatrix, Term, LinearOperatorFamily, pow0
def run(n, dens, leaf, seed=0, grid=None):
    os.environ['WAE_LU_LEAF']=str(leaf)
    rng=np.random.default_rng(seed)
    if grid:
        nx,ny,nz=grid
        def lap(k): return sp.diags([-1,2.2,-1],[-1,0,1],shape=(k,k))
        A=sp.kronsum(sp.kronsum(lap(nx),lap(ny)),lap(nz)).tocsc().astype(complex)
        A=A+1j*sp.diags(rng.random(A.shape[0]))
        n=A.shape[0]
    else:
        A=sp.random(n,n,density=dens,random_state=seed,format='csc').astype(complex)
        A=A+1j*sp.random(n,n,density=dens,random_state=seed+1,format='csc')
        A=A+sp.diags(np.full(n,4.0+1j))
    A=sp.csc_matrix(A); A.sort_indices()
    L=LinearOperatorFamily(['w'],[0.0])
    L.push(Term(DeviceMatrix.from_scipy(A),(pow0,),(('w',),),'','A'))
    op=L(1.0)
    dev=L.device(); ctx=dev.ctx
    op.materialize(0); lid=dev.lu(); ctx.lu_factor(lid,0)
    B=rng.standard_normal((n,2))+1j*rng.standard_normal((n,2))
    errs=[]
    for trans,Aop in ((0,A),(1,A.T),(2,A.conj().T)):
        X=ctx.lu_solve(lid,B,trans=trans)
        Xr=spla.spsolve(sp.csc_matrix(Aop),B)
        errs.append(np.abs(X-Xr).max()/np.abs(Xr).max())
    print(f"n={n} leaf={leaf} grid={grid} nnzLU={dev.lu_nnz} errs={['%.1e'%e for e in errs]}")
run(20,0.3,1000)      # single supernode, one block
run(50,0.2,1000)      # single supernode, 2 blocks
run(200,0.05,1000)    # single supernode, 7 blocks
run(200,0.05,16)
run(0,0,8,grid=(6,6,1))
run(0,0,8,grid=(12,12,1))
run(0,0,16,grid=(10,10,10))
run(0,0,64,grid=(16,16,16))
