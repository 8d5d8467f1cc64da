# A/B of the backtrace fix: the regression tests must FAIL on the library built from the previous commit
# (tools/_ab/libxroute_b200_old.so -- build it here first, it is not kept:
#    mkdir -p /tmp/old tools/_ab && git archive <commit before ecdee9d> xroute_env_b200/csrc include | tar -x -C /tmp/old &&
#    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC \
#         -o tools/_ab/libxroute_b200_old.so /tmp/old/xroute_env_b200/csrc/xr_api.cu)
# and pass on the current one; then the fuzz
# campaign that found the case, the whole GPU suite and a bench line.
mkdir -p gpurun_out
XROUTE_B200_LIB=$PWD/tools/_ab/libxroute_b200_old.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "own_walk" > gpurun_out/walk_old.log 2>&1; echo "old lib rc=$? (expected 1)"
tail -3 gpurun_out/walk_old.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "own_walk or path_capacity" > gpurun_out/walk_new.log 2>&1; echo "new lib rc=$?"
tail -3 gpurun_out/walk_new.log
timeout 300 python tools/repro_fuzz.py > gpurun_out/repro_fuzz.log 2>&1; tail -5 gpurun_out/repro_fuzz.log
timeout 300 python tools/fuzz_parity.py 200 7 > gpurun_out/fuzz7.log 2>&1; tail -c 600 gpurun_out/fuzz7.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -c 1500 gpurun_out/bench_r1n.json
