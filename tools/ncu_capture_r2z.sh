tag=r2z
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > gpurun_out/${tag}_${name}_plain.log 2>&1 &&
  timeout 500 $NCU -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${tag}_${name} "$@" > gpurun_out/${tag}_${name}_ncu.log 2>&1
  echo "$name rc=$?"; tail -1 gpurun_out/${tag}_${name}_plain.log
}
cap dense_bwd tc_bwd_kernel 4 2 python tools/prof_case.py dense_bwd 1 8
cap circ_fwd tc_band_kernel 2 1 python tools/prof_case.py circ_fwd 1 64
cap circ_bwd tc_bwd_kernel 4 2 python tools/prof_case.py circ_bwd 1 32
cap win3d_bwd tc_win_bwd_kernel 2 1 python tools/prof_case.py win3d_bwd 1 2
ls -la gpurun_out/${tag}_*.ncu-rep
