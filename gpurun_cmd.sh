mkdir -p gpurun_out
timeout 300 python tools/repro_fuzz2.py > gpurun_out/repro.log 2>&1; tail -5 gpurun_out/repro.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "frontier or capacity or small_grids" > gpurun_out/t1.log 2>&1; tail -3 gpurun_out/t1.log
timeout 600 python tools/fuzz_parity.py 200 12 ties > gpurun_out/fuzz2.log 2>&1; tail -2 gpurun_out/fuzz2.log | cut -c1-600
timeout 600 python tools/fuzz_parity.py 150 13 > gpurun_out/fuzz3.log 2>&1; tail -2 gpurun_out/fuzz3.log | cut -c1-600
timeout 300 python tools/quick_bench.py SYN-256 64 32 2 0 2>&1 | tail -2 | head -1
