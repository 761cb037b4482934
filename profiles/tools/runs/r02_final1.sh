set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest.log; tail -3 $O/r02k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02k_smoke.log 2>&1; tail -2 $O/r02k_smoke.log
python bench.py > $O/r02k_bench1.json 2> $O/r02k_bench1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02k_ref.json 2> $O/r02k_ref.err; echo "ref rc=$?"
B="python bench.py --no-cpu-baseline --no-also --no-solver"
$B --workload c2 --op C > $O/r02k_c2C.json 2>> $O/r02k.err
$B --workload c4 > $O/r02k_c4T.json 2>> $O/r02k.err
$B --workload c4 --op N > $O/r02k_c4N.json 2>> $O/r02k.err
$B --workload c3 --op T > $O/r02k_c3T.json 2>> $O/r02k.err
$B --workload c1 > $O/r02k_c1.json 2>> $O/r02k.err
$B --workload c5 > $O/r02k_c5.json 2>> $O/r02k.err
B2="python bench.py --no-cpu-baseline --no-also --no-solver --steps 3 --warmup 3"
run_full() {   # name, kernel regex, bench args...
    local name=$1 rx=$2; shift 2
    ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o $O/$name $B2 "$@" > $O/${name}_ncu.log 2>&1
    echo "$name: ncu rc=$?"
}
run_full r02k_c3N_stream_warp_full stream_warp --workload c3
run_full r02k_c3T_stream_warp_full stream_warp --workload c3 --op T
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02k_c3_launches.csv $B2 --workload c3 > $O/r02k_c3_launches_ncu.log 2>&1
python - <<PY
import json
for f in ["bench1","ref","c2C","c4T","c4N","c3T","c1","c5"]:
    try:
        d=json.loads(open("gpurun_out/r02k_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("frac"), (d.get("parity") or {}).get("rel_err"), (d.get("e2e") or {}).get("ms_per_step"))
        if f=="bench1":
            for a in d.get("also",[]): print("   also", a["workload"][:10], a["ms_per_step"], a["roofline"]["frac"], a["parity"]["rel_err"])
            print("   solver", d.get("solver")); print("   cpu", d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
