import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_stack import _stack, _run, _oracle
from astrophotography_b200 import kernels
def show(n, case, quantise, out_f64):
    method, k_lo, k_hi, maxiters, cen, dev = case
    st = _stack(n, (9, 70), seed=2, quantise=quantise)
    exp = _oracle(st, method, k_lo, k_hi, maxiters, cen, dev)
    got = _run(torch, st, method=method, k_lo=k_lo, k_hi=k_hi, maxiters=maxiters, cen=cen, dev=dev, out_f64=out_f64)
    badn = got["nrej"].astype(np.int64) != exp["nrej"]
    g = got["data"].astype(np.float64); e = exp["data"]
    with np.errstate(invalid='ignore'):
        rel = np.abs(g-e)/np.maximum(np.abs(e),1.0)
    rel[np.isnan(g)&np.isnan(e)] = 0
    rel[np.isinf(e)&(g==e)] = 0
    print(n, case, 'q',quantise,'f64',out_f64, 'nrej mismatches', badn.sum(), 'max rel', np.nanmax(rel), 'nan mismatch', (np.isnan(g)!=np.isnan(e)).sum())
    idx = np.argwhere(badn | (rel>1e-6) | (np.isnan(g)!=np.isnan(e)))[:6]
    for (r,c) in idx:
        print('   px',r,c,'got',g[r,c],got['nrej'][r,c],'exp',e[r,c],exp['nrej'][r,c], 'vals', np.sort(st[:,r,c])[:4], np.sort(st[:,r,c])[-3:])
show(3, ("average",5.0,5.0,1,"median","mad_std"), False, False)
show(100, ("average",3.0,3.0,1,"median","mad_std"), False, False)
show(100, ("average",3.0,3.0,1,"median","mad_std"), False, True)
show(100, ("average",3.0,3.0,5,"mean","std"), False, False)
show(100, ("average",3.0,3.0,5,"mean","std"), False, True)
show(100, ("average",3.0,3.0,5,"mean","std"), True, True)
show(5, ("average",3.0,3.0,0,"mean","std"), False, True)
show(16, ("average",3.0,3.0,0,"mean","std"), False, True)
