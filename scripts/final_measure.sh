#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): bench lines, ncu launch list, per-kernel metrics, full capture of the dominant kernel.
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench_N1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-forward --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_rotate_b8_metrics.csv python scripts/prof_rotate.py 16 8 28 2 > gpurun_out/${TAG}_ncu_rotate.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ntt_chunk_kernel -s 1 -c 1 -o gpurun_out/${TAG}_ntt_chunk_full python scripts/prof_rotate.py 16 8 28 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_ntt_chunk_full.ncu-rep --page details > gpurun_out/${TAG}_ncu_full_ntt_chunk_kernel.txt 2>/dev/null
python scripts/forward_profile.py > gpurun_out/${TAG}_forward_profile.txt 2>&1
python scripts/prof_rotsum.py 10 64 3 >> gpurun_out/${TAG}_forward_profile.txt 2>&1
tail -c 600 gpurun_out/${TAG}_bench_N1.json
