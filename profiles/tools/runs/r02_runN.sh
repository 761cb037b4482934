# usage: r02_runN.sh N tag
N=$1; TAG=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench$N.json").read().strip().splitlines()[-1])
print("c2", d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["rel_err"], d["roofline"].get("kernel_ms"))
for a in d.get("also", []):
    print(a["workload"][:12], a["ms_per_step"], a["e2e"]["ms_per_step"], a["parity"]["rel_err"], a.get("parallelism","")[:70], a["roofline"].get("kernel_ms"))
print(d.get("solver"))
PY
