# Round-2 end-of-round evidence run (one GPU): GPU tests, smoke, bench (both arms), the reference's benchmark tables,
# softmax bench, the ncu launch list of the bench command and one --set full capture of the headline kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_pytest_gpu.log; cat gpurun_out/r2z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2>gpurun_out/r2z_bench.err | tail -1 > gpurun_out/r2z_bench_n1.json; cut -c1-300 gpurun_out/r2z_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2z_bench_reference.json; cut -c1-200 gpurun_out/r2z_bench_reference.json
timeout 300 python tools/bench_compare.py --reps 20 > gpurun_out/r2z_bench_compare.jsonl 2>gpurun_out/r2z_bench_compare.err; wc -l gpurun_out/r2z_bench_compare.jsonl
timeout 300 python tools/bench_softmax.py --reps 20 > gpurun_out/r2z_softmax_bench.jsonl 2>gpurun_out/r2z_softmax_bench.err; tail -8 gpurun_out/r2z_softmax_bench.jsonl
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --batch 128 > gpurun_out/r2z_fwd_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fwd_kernel -s 3 -c 1 -f -o gpurun_out/r2z_dense_fwd python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --batch 128 > gpurun_out/r2z_fwd_ncu.log 2>&1; echo ncu rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2z_bench_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2z_bench_ncu.log 2>&1; echo launches rc=$?
