timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log; tail -5 gpurun_out/r02x_pytest.log
B="python bench.py --no-cpu-baseline"
$B --workload c4 --op N > gpurun_out/r02x_c4N.json 2> gpurun_out/r02x.err
$B --workload c4 --op N --variant 1 > gpurun_out/r02x_c4N_gather.json 2>> gpurun_out/r02x.err
$B --workload c4 > gpurun_out/r02x_c4T.json 2>> gpurun_out/r02x.err
$B --workload c1 > gpurun_out/r02x_c1.json 2>> gpurun_out/r02x.err
$B --workload c1 --warm-l2 > gpurun_out/r02x_c1_warm.json 2>> gpurun_out/r02x.err
$B --workload c3 --op T > gpurun_out/r02x_c3T.json 2>> gpurun_out/r02x.err
tail -3 gpurun_out/r02x.err
python - <<PY
import json
for f in ["c4N","c4N_gather","c4T","c1","c1_warm","c3T"]:
    try:
        d=json.loads(open("gpurun_out/r02x_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), d["roofline"]["frac"], d["parity"]["rel_err"], d["config"].get("plan"))
    except Exception as e: print(f, "ERR", e)
PY
