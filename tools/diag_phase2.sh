# phase timing of the dual-layout kernel (rank 0 of its clusters; the band kernel reports only for C >= 99 here)
# names: ph_compact_y=y phase, ph_sweep_y=wait after y, ph_compact_x=x phase, ph_sweep_x=z phase, ph_sweep_z=closing wait,
#        ph_halo_pull=dirty columns, ph_end_sync=dirty rows
set -e
cp xroute_env_b200/libxroute_b200.so /tmp/keep.so
XR_NVCC_EXTRA="-DWIN_PHASE_TIMING -DWIN_PHASE_MINC=99" python -m xroute_env_b200.build --force >/dev/null 2>&1
python tools/diag_route.py 2>&1 | head -1
cp /tmp/keep.so xroute_env_b200/libxroute_b200.so
