# full bench line + the reference arm (CPU oracle on every host core), as the driver runs them
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("ours", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["ispd18_test1"]["value"],
      d["ispd18_test1"]["x_over_cpu_all_cores"], d["cpu_baseline"]["value"], d["full_obs_rebuild"]["value"])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final_ref.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["cpu_baseline"])
PY
