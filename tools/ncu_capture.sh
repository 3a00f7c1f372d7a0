#!/bin/bash
# ncu captures of the hot kernels (run on a GPU box: gpurun -- 'bash tools/ncu_capture.sh r1s').
# Every case runs once without ncu first (must exit 0), then under `ncu --set full` for the named kernel.
# Output: gpurun_out/<tag>_<case>.ncu-rep; read here with `ncu -i ... --page raw --csv` (tools/ncu_summary.py).
tag=${1:-cap}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > gpurun_out/${tag}_${name}_plain.log 2>&1 &&
  timeout 600 $NCU -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${tag}_${name} "$@" > gpurun_out/${tag}_${name}_ncu.log 2>&1
  echo "$name rc=$?"; tail -1 gpurun_out/${tag}_${name}_plain.log
}
cap dense_fwd tc_fwd_kernel 2 1 python tools/prof_case.py dense_fwd 1 32
cap dense_bwd tc_bwd_kernel 4 2 python tools/prof_case.py dense_bwd 1 8
cap circ_fwd tc_fwd_kernel 2 1 python tools/prof_case.py circ_fwd 1 64
cap circ_bwd tc_bwd_kernel 4 2 python tools/prof_case.py circ_bwd 1 32
cap win3d_fwd tc_win_fwd_kernel 2 1 python tools/prof_case.py win3d_fwd 1 2
cap win3d_bwd tc_win_bwd_kernel 2 1 python tools/prof_case.py win3d_bwd 1 2
# launch list of the bench command (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_bench_ncu.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/${tag}_bench_plain.log | cut -c1-300
ls -la gpurun_out/${tag}_*.ncu-rep
