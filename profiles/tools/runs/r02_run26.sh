timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest.log; tail -3 gpurun_out/r02y_pytest.log
B="python bench.py --no-cpu-baseline --workload c3"
for W in 4096 5120 3072; do
  BSM_TUNE_WCHUNK=$W $B --op T > gpurun_out/r02y_c3T_$W.json 2>> gpurun_out/r02y.err
  BSM_TUNE_WCHUNK=$W $B > gpurun_out/r02y_c3N_$W.json 2>> gpurun_out/r02y.err
done
tail -3 gpurun_out/r02y.err
python - <<PY
import json
for f in ["c3T_4096","c3N_4096","c3T_5120","c3N_5120","c3T_3072","c3N_3072"]:
    try:
        d=json.loads(open("gpurun_out/r02y_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), d["roofline"]["frac"], d["parity"]["rel_err"], d["config"]["plan"]["warp_chunks"])
    except Exception as e: print(f, "ERR", e)
PY
