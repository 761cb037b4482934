timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_pytest.log
B="python bench.py --no-cpu-baseline"
$B --workload c3 --op T > gpurun_out/r02d_c3T.json 2> gpurun_out/r02d.err
$B --workload c3 > gpurun_out/r02d_c3N.json 2>> gpurun_out/r02d.err
$B --workload c1 > gpurun_out/r02d_c1.json 2>> gpurun_out/r02d.err
$B --workload c1 --op T > gpurun_out/r02d_c1T.json 2>> gpurun_out/r02d.err
tail -3 gpurun_out/r02d.err
python - <<PY
import json
for f in ["c3T","c3N","c1","c1T"]:
    try:
        d=json.loads(open("gpurun_out/r02d_%s.json"%f).read().strip().splitlines()[-1])
        pl=d["config"].get("plan",{})
        print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], pl.get("warp_chunks"))
    except Exception as e: print(f, "ERR", e)
PY
