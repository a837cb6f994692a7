#!/bin/bash
# instruction mix of the set-up + epilogue: the kernel at t = 2 (one time step)
ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64_op_dfma.sum \
  --clock-control none -k regex:fd1d_iw -c 2 --csv --log-file gpurun_out/r2v_setup_ncu.csv python tools/variant_probe.py 1024 2 32768 237 > /dev/null 2>&1
cat gpurun_out/r2v_setup_ncu.csv | tail -32
