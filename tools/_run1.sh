export PYTHONUNBUFFERED=1
echo "== wbox=256"
MANSY_MTIO_WBOX=256 timeout 60 python tools/mtio_bench.py 16384 5 2>&1 | grep -E "tf32 max|tf32 batch 16384:|serialised" | head -3
MANSY_MTIO_WBOX=256 timeout 120 python -m pytest tests/test_gpu_mtio.py -x -q 2>&1 | tail -3
