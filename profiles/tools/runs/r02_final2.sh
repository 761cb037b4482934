set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02p2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p2_pytest.log; tail -3 $O/r02p2_pytest.log
B="python bench.py --no-cpu-baseline"
$B --workload c1 --warm-l2 > $O/r02p2_c1_warm.json 2> $O/r02p2.err
$B --workload c1 > $O/r02p2_c1.json 2>> $O/r02p2.err
$B --workload c3 > $O/r02p2_c3N.json 2>> $O/r02p2.err
python - <<PY
import json
for f in ["c1_warm","c1","c3N"]:
    d=json.loads(open("gpurun_out/r02p2_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"].get("kernel_ms"), round(d["roofline"]["frac"],3), d["parity"]["rel_err"], d["config"].get("l2"))
PY
