# CPU model: how many polishing sweeps until bitwise convergence, for typical and stiff PDEs
import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R+'/oracle', R+'/kwinto-cuda_b200', R+'/tests']
import numpy as np, pyoracle, scheme_model
from kwfd1d.synthetic import synthetic_options
o = pyoracle.Oracle()
def sweeps(opt, x, t, nodes_per=32):
    tt,k,z,r,q,s,e,w = (opt[f] for f in ("t","k","z","r","q","s","e","w"))
    xs = o.x_grid(float(z), float(tt), x)
    bl,b,bu = scheme_model.coefficients(xs, -r, r-q-z*z/2, z*z/2, tt/(t-1))
    beta = scheme_model.pivots_serial(bl,b,bu)
    mo = scheme_model.pivots_moebius(bl,b,bu,8)
    relerr = np.max(np.abs(mo-beta)/np.abs(beta))
    L = x//nodes_per
    pin = np.full(L, np.inf); pin[1:] = mo[nodes_per-1:-1:nodes_per][:L-1]
    for sweep in range(200):
        out = np.empty(L)
        for l in range(L):
            prev = pin[l]
            for j in range(l*nodes_per,(l+1)*nodes_per):
                gam = (bu[j-1] if j>0 else 0.0)/prev
                prev = b[j]-bl[j]*gam
            out[l]=prev
        pnew = np.full(L, np.inf); pnew[1:]=out[:-1]
        if np.array_equal(pnew.view(np.int64), pin.view(np.int64)): return sweep+1, relerr
        pin=pnew
    return 200, relerr
opts = synthetic_options(8, 42)
print("1024x1024, 32 nodes/lane:", [sweeps(opts[i],1024,1024) for i in range(8)])
opts = synthetic_options(8, 32, call_every=2)
print("4096x32, 8 nodes/chunk:", [sweeps(opts[i],4096,32,8) for i in range(4)])
print("1024x16, 32 nodes/lane:", [sweeps(opts[i],1024,16,32) for i in range(4)])
