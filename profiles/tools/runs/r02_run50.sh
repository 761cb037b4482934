timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4 or tall or fuzz or odd" > gpurun_out/r02t2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t2_pytest.log; tail -3 gpurun_out/r02t2_pytest.log
python bench.py --no-cpu-baseline --workload c4 > gpurun_out/r02t2_c4T.json 2> gpurun_out/r02t2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02t2_c4T.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], d["config"]["plan"]["slices"]["sym_fused_tma_kernel"])
PY
