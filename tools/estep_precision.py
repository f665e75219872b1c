import numpy as np, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')  # run from the repo root
from conftest import load_golden
from oracle import gmm as og
g = load_golden('gmm','c1')
X = g['z'].astype(np.float64)
it=4
w, mu, cov, P = g['it_weights'][it], g['it_means'][it], g['it_covariances'][it], g['it_pchol'][it]
lp64 = og.log_gaussian_prob(X, mu, P) + np.log(w)
f32=np.float32
def estep(mu_mode, u_mode, acc_mode):
    K,d = mu.shape
    out = np.empty((X.shape[0],K))
    for k in range(K):
        if mu_mode=='f32': diff = (X.astype(f32) - mu[k].astype(f32)).astype(f32)
        elif mu_mode=='hilo':
            hi = mu[k].astype(f32); lo=(mu[k]-hi.astype(np.float64)).astype(f32)
            diff = ((X.astype(f32)-hi).astype(f32) - lo).astype(f32)
        else: diff = X-mu[k]
        U = P[k].astype(f32) if u_mode=='f32' else P[k]
        if acc_mode=='f32':
            y = np.zeros((X.shape[0],d),dtype=f32)
            for b in range(d):
                acc=np.zeros(X.shape[0],dtype=f32)
                for c in range(b+1):
                    acc = (acc + (diff[:,c].astype(f32)*f32(U[c,b])).astype(f32)).astype(f32)  # not fma but close
                y[:,b]=acc
            m = (y.astype(f32)**2).sum(1,dtype=f32)
        else:
            y = diff.astype(np.float64) @ np.asarray(U,dtype=np.float64)
            m = (y**2).sum(1)
        out[:,k] = -0.5*(d*np.log(2*np.pi)+m.astype(np.float64)) + np.sum(np.log(np.diag(P[k]))) + np.log(w[k])
    return out
def report(name, lp):
    # error on the dominant component + on lse
    lse64 = np.logaddexp.reduce(lp64,axis=1); lse=np.logaddexp.reduce(lp,axis=1)
    r64=np.exp(lp64-lse64[:,None]); r=np.exp(lp-lse[:,None])
    print(f"{name:28s} max|dlse| {np.abs(lse-lse64).max():.2e} mean {np.abs(lse-lse64).mean():.2e}  max|dr| {np.abs(r-r64).max():.2e}")
for mm in ['f32','hilo','f64']:
    for um in ['f32','f64']:
        for am in ['f32','f64']:
            report(f"mu={mm} U={um} acc={am}", estep(mm,um,am))
print('cond', [f"{np.linalg.cond(c):.1e}" for c in cov])
print('max |U|', np.abs(P).max(), 'mu mag', np.abs(mu).max())
