timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
B="python bench.py --no-cpu-baseline"
$B --workload c4 --op N > gpurun_out/r02z_c4N.json 2> gpurun_out/r02z.err
$B --workload c4 > gpurun_out/r02z_c4T.json 2>> gpurun_out/r02z.err
$B --workload c3 --op T > gpurun_out/r02z_c3T.json 2>> gpurun_out/r02z.err
$B --workload c3 > gpurun_out/r02z_c3N.json 2>> gpurun_out/r02z.err
$B --workload c1 > gpurun_out/r02z_c1.json 2>> gpurun_out/r02z.err
$B --workload c2 --no-also --no-solver > gpurun_out/r02z_c2.json 2>> gpurun_out/r02z.err
tail -3 gpurun_out/r02z.err
python - <<PY
import json
    try:
        d=json.loads(open("gpurun_out/r02z_%s.json"%f).read().strip().splitlines()[-1])
        pl=d["config"].get("plan",{})
        print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), d["roofline"]["frac"], d["parity"]["rel_err"], pl.get("slices"), pl.get("warp_chunks"))
    except Exception as e: print(f, "ERR", e)
PY
