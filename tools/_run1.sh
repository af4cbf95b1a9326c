export PYTHONUNBUFFERED=1
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01z_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# --set full captures: MTIO GEMM tiles and attention (decode steady state), expert kernel
MTIO_BENCH_PROFILE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtio_gemm --launch-skip 250 --launch-count 6 -o gpurun_out/r01z_mtio_gemm python tools/mtio_bench.py 16384 1 > /dev/null 2>&1
MTIO_BENCH_PROFILE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtio_attn --launch-skip 84 --launch-count 2 -o gpurun_out/r01z_mtio_attn python tools/mtio_bench.py 16384 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:expert_mpc --launch-skip 12 --launch-count 1 -o gpurun_out/r01z_expert python tools/expert_bench.py 4096 > /dev/null 2>&1
ls -la gpurun_out | tail -8
