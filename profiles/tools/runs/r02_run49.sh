timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s2_pytest.log; tail -3 gpurun_out/r02s2_pytest.log
B="python bench.py --no-cpu-baseline --workload c4"
$B > gpurun_out/r02s2_c4T.json 2> gpurun_out/r02s2.err
$B --op N > gpurun_out/r02s2_c4N.json 2>> gpurun_out/r02s2.err
python - <<PY
import json
for f in ["c4T","c4N"]:
    d=json.loads(open("gpurun_out/r02s2_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], d["config"]["plan"]["slices"]["sym_fused_tma_kernel"], d["e2e"]["ms_per_step"])
PY
