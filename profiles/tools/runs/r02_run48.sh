B="python bench.py --no-cpu-baseline --workload c4"
for W in 524288 1048576 2097152 262144; do
  BSM_TUNE_WORK_TARGET=$W $B > gpurun_out/r02r2_c4T_$W.json 2> gpurun_out/r02r2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02r2_c4T_$W.json").read().strip().splitlines()[-1])
print($W, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], d["config"]["plan"]["slices"]["sym_fused_tma_kernel"])
PY
done
