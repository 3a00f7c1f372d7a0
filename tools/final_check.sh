set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1w_pytest_gpu.log; cat gpurun_out/r1w_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2>gpurun_out/r1w_bench.err | tail -1 > gpurun_out/r1w_bench_n1.json; cut -c1-400 gpurun_out/r1w_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1w_bench_reference.json; cut -c1-300 gpurun_out/r1w_bench_reference.json
python tools/prof_case.py circ_fwd 1 64 > gpurun_out/r1w_circ_fwd_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_band_kernel -s 2 -c 1 -f -o gpurun_out/r1w_circ_fwd python tools/prof_case.py circ_fwd 1 64 > gpurun_out/r1w_circ_fwd_ncu.log 2>&1; echo ncu rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1w_bench_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r1w_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1w_bench_ncu.log 2>&1; echo launches rc=$?
