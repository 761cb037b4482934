for W in 0 3584 2688; do
  if [ $W -gt 0 ]; then export BSM_TUNE_WCHUNK=$W; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --workload c3 --no-cpu-baseline > gpurun_out/r02g_c3_4_$W.json 2> gpurun_out/r02g.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02g_c3_4_$W.json").read().strip().splitlines()[-1])
print($W, d["ms_per_step"], d["roofline"].get("kernel_ms"), d["parity"]["rel_err"], d["config"]["plan"]["warp_chunks"])
PY
done
