python bench.py > gpurun_out/r02z9_bench1_final.json 2> gpurun_out/r02z9_bench1_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02z9_bench1_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["parity"]["rel_err"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"])
for a in d.get("also",[]): print("   also", a["workload"][:10], a["ms_per_step"], a["roofline"]["frac"], a["parity"]["rel_err"])
print(d.get("cpu_baseline"))
PY
