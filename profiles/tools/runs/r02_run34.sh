timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
for W in 0 1; do
  export BSM_TUNE_NO_XFIFO=$W
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --workload c3 --no-cpu-baseline > gpurun_out/r02h_c3_4_noxf$W.json 2> gpurun_out/r02h.err; echo "rc=$?"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --workload c3 --op T --no-cpu-baseline > gpurun_out/r02h_c3T_4_noxf$W.json 2>> gpurun_out/r02h.err; echo "rc=$?"
  python - <<PY
import json
for f in ["c3","c3T"]:
    d=json.loads(open("gpurun_out/r02h_%s_4_noxf$W.json"%f).read().strip().splitlines()[-1])
    print(f, "noxf=$W", d["ms_per_step"], d["roofline"].get("kernel_ms"), d["parity"]["rel_err"], d["config"]["plan"]["warp_chunks"])
PY
done
tail -3 gpurun_out/r02h.err
