timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or c1_shape or vbcrs" > gpurun_out/r02w2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w2_pytest.log; tail -2 gpurun_out/r02w2_pytest.log
B="python bench.py --no-cpu-baseline --workload c3"
$B --op T > gpurun_out/r02w2_c3T.json 2> gpurun_out/r02w2.err
$B > gpurun_out/r02w2_c3N.json 2>> gpurun_out/r02w2.err
python - <<PY
import json
for f in ["c3T","c3N"]:
    d=json.loads(open("gpurun_out/r02w2_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], d["config"]["plan"]["warp_chunks"])
PY
