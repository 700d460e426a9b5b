timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest_final.log; cat gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
