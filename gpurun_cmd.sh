mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_nccl_stats.py -q -x > gpurun_out/tnccl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tnccl.log
tail -25 gpurun_out/tnccl.log
