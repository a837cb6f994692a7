#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small invocations of every kernel the shipped dispatch can reach:
# tools/gpu_sanitize_r2.sh
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_sanitize.log) 2>&1
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import numpy as np, kwfd1d
from kwfd1d.synthetic import synthetic_options
def run(x,t,n,variant=0,prec="f64",mode="FD1D-GPU",**kw):
    cfg=kwfd1d.Config(PRICER=mode); cfg.set("FD1D.T_GRID_SIZE",t); cfg.set("FD1D.X_GRID_SIZE",x)
    cfg.set("FD1D.GPU.VARIANT",variant); cfg.set("FD1D.GPU.PRECISION",prec)
    for k,v in kw.items(): cfg.set(k,v)
    err,p=kwfd1d.PricerFactory.create(cfg); assert err=="",err
    o=synthetic_options(n,5,european_every=4,call_every=3); o=np.concatenate([o,o[:n//5]])
    err,got=p.price(o); assert err=="",err
    print("ok",mode,x,t,n,p.info()["variant"],float(got.sum()))
run(1024,12,44,237); run(700,12,41,237); run(1024,12,24,201); run(512,12,41,138); run(300,12,43,138); run(512,12,24,101)
run(256,12,41,38); run(70,12,43,38); run(256,12,40,1); run(2048,10,24,336); run(1100,10,9,336); run(4096,8,12,436); run(3000,8,5,436); run(2048,8,6,301); run(4096,8,4,401)
run(1024,12,41,1237,"f32"); run(512,12,41,1138,"f32"); run(256,12,43,1038,"f32"); run(512,12,40,1101,"f32"); run(1024,12,24,1201,"f32")
run(1024,12,42,0,mode="FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":4}); run(512,12,43,0,mode="FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":4})
run(256,12,41,0,mode="FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":4}); run(2048,8,9,0,mode="FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":4})
run(512,10,1100,0)  # device-side compression, both batch-size classes launched (class split)
def long_chain(x,t,mode="FD1D-GPU",**kw):
    cfg=kwfd1d.Config(PRICER=mode); cfg.set("FD1D.T_GRID_SIZE",t); cfg.set("FD1D.X_GRID_SIZE",x)
    for k,v in kw.items(): cfg.set(k,v)
    err,p=kwfd1d.PricerFactory.create(cfg); assert err=="",err
    o=synthetic_options(1000,6,european_every=4,call_every=3); e=np.repeat(o[5:6],700); e["k"]*=np.linspace(0.8,1.2,700); o=np.concatenate([o,e])
    err,got=p.price(o); assert err=="",err
    print("ok long chain",mode,x,t,p.info()["variant"],p.info()["long_chains"],float(got.sum()))
long_chain(1024,8); long_chain(512,8); long_chain(256,8); long_chain(1024,8,"FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":4})
run(1024,12,42,0,mode="FD1D-BS-GPU",**{"FD1D.GPU.BS_FUSED":1}); run(1024,8,40,0,**{"FD1D.GPU.LAYOUT":"soa"})
run(1024,12,44,0,**{"FD1D.GPU.DEVICES":"0,0"})
PY
for tool in memcheck racecheck; do
  echo "== $tool"; timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -vE "^$" | tail -30
done
