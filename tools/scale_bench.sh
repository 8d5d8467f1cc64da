# weak-scaling lines as the driver launches them (run under: gpurun --gpus 8 -- bash tools/scale_bench.sh)
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 64 --warmup 4 > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_scale_n$n.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["ms_per_step_by_rank"], d["clocks"])
PY
done
