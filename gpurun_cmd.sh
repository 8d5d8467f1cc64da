mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "frontier or small_grids or backtrace or long_paths or metrics_by or tie_heavy or nonuniform or reset" > gpurun_out/t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t1.log
tail -5 gpurun_out/t1.log
timeout 600 python tools/quick_bench.py SYN-256 64 32 2 0 > gpurun_out/qb.log 2>&1
tail -3 gpurun_out/qb.log
timeout 600 python tools/quick_bench.py T1-7x7 512 32 1 0 > gpurun_out/qb_t17.log 2>&1
tail -3 gpurun_out/qb_t17.log
