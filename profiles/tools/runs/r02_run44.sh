timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r02n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n2_pytest.log; tail -5 gpurun_out/r02n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-also > gpurun_out/r02n2_bench2.json 2> gpurun_out/r02n2_bench2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02n2_bench2.json").read().strip().splitlines()[-1])
print("c2", d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["rel_err"])
print(d.get("solver"))
PY
