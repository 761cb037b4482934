timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02q2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q2_pytest.log; tail -3 gpurun_out/r02q2_pytest.log
B="python bench.py --no-cpu-baseline"
$B --workload c3 --op T > gpurun_out/r02q2_c3T.json 2> gpurun_out/r02q2.err
$B --workload c3 --op T > gpurun_out/r02q2_c3T_b.json 2>> gpurun_out/r02q2.err
python - <<PY
import json
for f in ["c3T","c3T_b"]:
    d=json.loads(open("gpurun_out/r02q2_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"])
PY
