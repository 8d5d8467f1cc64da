# first-connection vs later-connection cost of the window kernel (rank-0 CTAs of clusters >= MINC)
# names: ph_compact_y=first iters, ph_sweep_y=first relax cycles, ph_compact_x=later iters, ph_sweep_x=later relax cycles,
#        ph_sweep_z=connections, ph_halo_pull=post-relax cycles (target/exit/backtrace/commit)
set -e
cp xroute_env_b200/libxroute_b200.so /tmp/keep.so
for f in "-DWIN_PHASE_MINC=8" "-DWIN_PHASE_MINC=1"; do
XR_NVCC_EXTRA="-DWIN_PHASE_TIMING -DWIN_PHASE_SPLIT $f" python -m xroute_env_b200.build --force >/dev/null 2>&1
echo "== $f" >> gpurun_out/diag_split.log
python tools/diag_route.py 2>&1 | head -1 >> gpurun_out/diag_split.log
done
cp /tmp/keep.so xroute_env_b200/libxroute_b200.so
cat gpurun_out/diag_split.log
