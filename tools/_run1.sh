export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_mtio.py -x -q 2>&1 | tail -5 > gpurun_out/mtio_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --workload mtio --no-cpu-baseline > gpurun_out/bench_mtio.json 2>> gpurun_out/bench.err
cat gpurun_out/mtio_tests.log; tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_mtio.json
