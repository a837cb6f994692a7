#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/p.py <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import kwfd1d
print(kwfd1d.dfma_probe(0))
PY
timeout 600 ncu --clock-control none -k regex:dfma_operand_kernel --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,gpu__time_duration.sum --csv --log-file gpurun_out/ncu_dfma_probe.csv python /tmp/p.py > gpurun_out/ncu_dfma_probe.log 2>&1
tail -3 gpurun_out/ncu_dfma_probe.log
