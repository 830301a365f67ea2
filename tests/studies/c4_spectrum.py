"""Study (numpy oracle, CPU, ~1 min): why C4 fits need ~30 damped Newton iterations.

For one C4 toy generated at ``suggested_init`` this prints the spectrum of the exact Hessian of 2NLL
(forward differences of the oracle's analytic gradient) relative to the Fisher matrix (the same
Hessian on the Asimov data of the current point) at the start and along a few damped Newton steps.
Result recorded in DESIGN.md section 5: at the toy's own truth the exact Hessian has 39 negative
eigenvalues (min -127, Fisher min +2.0) and eig(F^-1 H) spans -33 ... +43 with a third of the
spectrum outside [0.5, 2]: the term sum_b (1 - n_b / lambda_b) d2(lambda_b) of 1000 bins x 99
normalisation systematics acting on every sample of every channel is a random symmetric matrix of
the size of the Fisher matrix itself.  Not a test (not collected); lives under tests/ because it
runs the oracle.

    python tests/studies/c4_spectrum.py
"""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from pyhf_b200 import synth
from pyhf_b200.compile import compile_spec
from oracle import hf_oracle as O
from oracle import fit_proto as FP
hp=compile_spec(synth.spec_c4(), poi_name="mu"); pl=O.Plan(hp.arrays())
rng=np.random.default_rng(5)
x0=hp.init.copy()
lam=O.expected_data(pl, x0[None])[0]
data=lam.copy(); data[:hp.n_bins]=rng.poisson(lam[:hp.n_bins])
# gaussian aux fluctuate
ga=np.asarray(hp.g_aux); data[hp.n_bins+ga]=lam[hp.n_bins+ga]+rng.normal(0,1,len(ga))*np.asarray(hp.g_sigma)
free=~hp.fixed.astype(bool)
t=time.time(); v,g=O.logpdf_grad(pl,x0[None],data[None]); print('eval s',time.time()-t, -2*v)
def hess(x,d): return FP.fd_hessian(pl,x[None],d[None],free)[0]
t=time.time(); Hx=hess(x0,data); print('hessian s',time.time()-t)
A=O.expected_data(pl,x0[None])[0]; Fi=hess(x0,A)
idx=np.nonzero(free)[0]
def spec(H,F):
    Hf=H[np.ix_(idx,idx)]; Ff=F[np.ix_(idx,idx)]
    L=np.linalg.cholesky(Ff); Li=np.linalg.inv(L)
    M=Li@Hf@Li.T; w=np.linalg.eigvalsh(0.5*(M+M.T)); return w
w=spec(Hx,Fi); print('at init: eig(F^-1 H) min %.3f max %.3f; outside [0.5,2]: %d of %d; negative: %d'%(w.min(),w.max(),((w<0.5)|(w>2)).sum(),len(w),(w<0).sum()))
print('eig(H) min', np.linalg.eigvalsh(Hx[np.ix_(idx,idx)]).min(), 'eig(F) min', np.linalg.eigvalsh(Fi[np.ix_(idx,idx)]).min())
# a few damped Newton iterations with exact H to move toward the minimum
x=x0.copy(); f=-2*v[0]
for it in range(6):
    v,g=O.logpdf_grad(pl,x[None],data[None]); g=-2*g[0]; f=-2*v[0]
    H=hess(x,data); Hf=H[np.ix_(idx,idx)]
    tau=0.0
    while True:
        try:
            w_=np.linalg.eigvalsh(Hf+tau*np.eye(len(idx)))
            if w_.min()<=1e-8: raise np.linalg.LinAlgError
            d=-np.linalg.solve(Hf+tau*np.eye(len(idx)),g[idx])
            xt=x.copy(); xt[idx]=np.clip(x[idx]+d,hp.lo[idx],hp.hi[idx])
            ft=-2*O.logpdf(pl,xt[None],data[None])[0]
            if np.isfinite(ft) and ft<f: break
        except np.linalg.LinAlgError: pass
        tau=max(8*tau,1e-2)
    A=O.expected_data(pl,x[None])[0]; F=hess(x,A); w=spec(H,F)
    print('it %d f %.4f -> %.4f tau %.3g |d|max %.3f ; eig(F^-1H) min %.3f max %.3f outside[0.5,2] %d neg %d'%(it,f,ft,tau,np.abs(d).max(),w.min(),w.max(),((w<0.5)|(w>2)).sum(),(w<0).sum()))
    x=xt
