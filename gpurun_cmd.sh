mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu > gpurun_out/tgpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tgpu.log
tail -30 gpurun_out/tgpu.log
