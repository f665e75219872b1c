import numpy as np, torch, sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import load_golden, rel_err
from spectrogram_cube_clustering_b200 import ops
import test_gpu_parity as T
for case in ['c1','k16','relu']:
    g = load_golden('gmm', case)
    iters = g['it_lower_bound'].shape[0]
    hist, resp0, labels = T._gmm_run(ops, g, iters)
    print(case, 'resp0', np.max(np.abs(resp0-np.exp(g['log_resp0']))))
    for it,h in enumerate(hist):
        print(it, 'lb %.3e'%abs(h['lb']-g['it_lower_bound'][it]), 'w %.2e'%rel_err(h['weights'], g['it_weights'][it]), 'mu %.2e'%rel_err(h['means'], g['it_means'][it]),
          'cov %.2e'%T._cov_err(h['cov'], g['it_covariances'][it]), 'pchol %.2e'%T._cov_err(h['pchol'], g['it_pchol'][it]),
          'cond %.1e'%max(np.linalg.cond(c) for c in g['it_covariances'][it]))
    print('label mism', (labels != g['labels_after']).mean())
try:
    z = torch.zeros(0, 9, device="cuda"); mu = torch.randn(8, 9, device="cuda")
    print(ops.dec_assign(z, mu))
except Exception as e: print('ERR', e)
