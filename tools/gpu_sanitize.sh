#!/bin/bash
# compute-sanitizer over small invocations of every march kernel: tools/gpu_sanitize.sh <tag>
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_sanitize.log) 2>&1
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import numpy as np, kwfd1d
from kwfd1d.synthetic import synthetic_options
def run(x,t,n,variant=0,prec="f64"):
    cfg=kwfd1d.Config(PRICER="FD1D-GPU"); cfg.set("FD1D.T_GRID_SIZE",t); cfg.set("FD1D.X_GRID_SIZE",x)
    cfg.set("FD1D.GPU.VARIANT",variant); cfg.set("FD1D.GPU.PRECISION",prec)
    err,p=kwfd1d.PricerFactory.create(cfg); assert err=="",err
    o=synthetic_options(n,5,european_every=4,call_every=3); o=np.concatenate([o,o[:n//5]])
    err,got=p.price(o); assert err=="",err
    print("ok",x,t,n,p.info()["variant"],float(got.sum()))
def run_bs(x,t,n,fused):
    cfg=kwfd1d.Config(PRICER="FD1D-BS-GPU"); cfg.set("FD1D.T_GRID_SIZE",t); cfg.set("FD1D.X_GRID_SIZE",x)
    cfg.set("FD1D.GPU.BS_FUSED",fused)
    err,p=kwfd1d.PricerFactory.create(cfg); assert err=="",err
    o=synthetic_options(n,5,european_every=4,call_every=3); o=np.concatenate([o,o[:n//5]])
    err,got=p.price(o); assert err=="",err
    print("ok bs",x,t,n,p.info()["variant"],float(got.sum()))
if len(sys.argv) > 1 and sys.argv[1] == "bs":
    run_bs(1024,12,42,4); run_bs(1024,12,42,3); run_bs(1024,12,42,2); run_bs(512,12,42,4); run(1024,12,40,235)
    sys.exit(0)
run(1024,12,40,233); run(1024,12,40,241); run(1024,12,24,201); run(1024,12,24,221)
run(2048,10,24,331); run(4096,8,12,431); run(512,12,40,0); run(1024,12,24,0,"f32"); run(300,12,40,0)
run(512,12,40,133); run(300,12,40,133); run(1024,12,40,1233,"f32"); run(512,12,40,1133,"f32")
PY
for tool in memcheck racecheck; do
  echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py ${2:-} 2>&1 | grep -vE "^$" | tail -25
done
