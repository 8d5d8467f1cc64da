# phase timing of the window kernel.  Names in the CRIT runs: ph_compact_y=halo(max rank) ph_sweep_y=y(max) ph_compact_x=x(max)
# ph_sweep_x=z(max) ph_sweep_z=busiest rank y+x+z ph_halo_pull=idlest rank y+x+z ph_end_sync=rank-0 wait in the closing cluster barrier
set -e
cp xroute_env_b200/libxroute_b200.so /tmp/keep.so
for f in "-DWIN_PHASE_MINC=8" "-DWIN_PHASE_CRIT -DWIN_PHASE_MINC=8" "-DWIN_PHASE_CRIT -DWIN_PHASE_MINC=4"; do
XR_NVCC_EXTRA="-DWIN_PHASE_TIMING $f" python -m xroute_env_b200.build --force >/dev/null 2>&1
echo "== $f" >> gpurun_out/diag_phase.log
python tools/diag_route.py 2>&1 | head -1 >> gpurun_out/diag_phase.log
done
cp /tmp/keep.so xroute_env_b200/libxroute_b200.so
cat gpurun_out/diag_phase.log
