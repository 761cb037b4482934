set -u
O=gpurun_out
B2="python bench.py --no-cpu-baseline --no-also --steps 1 --warmup 3 --workload c2"
$B2 > $O/r02l_c2_plain.json 2> $O/r02l_c2_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02l_c2_solver_launches.csv $B2 > $O/r02l_c2_launches_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02l_c2_solver_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split("(")[0][-60:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])
for k,(n,t) in agg.items(): print(n, round(t/n/1000,2), "us avg", k)
PY
