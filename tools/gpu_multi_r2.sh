#!/bin/bash
# multi-GPU session: usage tools/gpu_multi_r2.sh N  (run under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_multi_${N}gpu.log) 2>&1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== multi-device C-ABI handle (one process, $N GPUs)"
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_device" 2>&1 | tail -3
echo "== bench default, $N ranks (weak + strong sub-record + C-ABI multi-device record)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -2 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2_bench_${N}gpu.json") if l.startswith("{")][-1]
print("weak value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4))
s = d.get("strong"); print("strong", {k: v for k, v in s.items() if k not in ("workload", "gather", "e2e")}, "e2e", round(s["e2e"]["value"]))
print("c_abi_multi_device", {k: v for k, v in d.get("c_abi_multi_device", {}).items() if k != "what"})
PY
echo "== one process, C-ABI multi-device handle"; timeout 600 python tools/multi_device_probe.py
echo "== bench --config 4, $N ranks"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
tail -2 gpurun_out/r2_bench_c4_${N}gpu.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2_bench_c4_${N}gpu.json") if l.startswith("{")][-1]
print("config 4 value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "gather_ms", round(d.get("gather_ms", 0), 3), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"], 2))
PY
