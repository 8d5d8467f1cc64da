set -e
cp xroute_env_b200/libxroute_b200.so /tmp/keep.so
XR_NVCC_EXTRA=-DWIN_PHASE_TIMING python -m xroute_env_b200.build --force >/dev/null
python tools/diag_route.py > gpurun_out/diag_phase.log 2>&1
cp /tmp/keep.so xroute_env_b200/libxroute_b200.so
tail -5 gpurun_out/diag_phase.log
