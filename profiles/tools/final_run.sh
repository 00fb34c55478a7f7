python -m pytest tests -m gpu -q > gpurun_out/s4_pytest_final.log 2>&1; tail -2 gpurun_out/s4_pytest_final.log
python bench.py > gpurun_out/s4_bench_final.log 2> gpurun_out/s4_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_ref_final.log 2>/dev/null
tail -c 300 gpurun_out/s4_ref_final.log
