timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log; tail -15 gpurun_out/r02e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02e_bench2.json 2> gpurun_out/r02e_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02e_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench2.json").read().strip().splitlines()[-1])
print("c2", d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["rel_err"])
for a in d.get("also", []):
    print(a["workload"][:12], a["ms_per_step"], a["e2e"]["ms_per_step"], a["parity"]["rel_err"], a.get("parallelism","")[:60])
print(d.get("solver"))
PY
