export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_mtio.py -x -q 2>&1 | tail -5 > gpurun_out/mtio_tests.log
timeout 200 python tools/mtio_bench.py 16384 5 > gpurun_out/mtio_bench.log 2>&1
MTIO_BENCH_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mtio_launches.csv python tools/mtio_bench.py 16384 1 > /dev/null 2>&1
cat gpurun_out/mtio_tests.log gpurun_out/mtio_bench.log
