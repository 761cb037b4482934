N=$1
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c3 --no-cpu-baseline > gpurun_out/r02j_c3_${N}_$tag.json 2> gpurun_out/r02j.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02j_c3_${N}_$tag.json").read().strip().splitlines()[-1])
print("$tag", d["ms_per_step"], d["roofline"].get("kernel_ms"), d["parity"]["rel_err"], d["config"]["plan"]["warp_chunks"])
PY
}
run bulk BSM_DUMMY=0
run elem BSM_TUNE_NO_XBULK_PEER=1
run bulk4096 BSM_TUNE_WCHUNK=4096
run elem4096 BSM_TUNE_NO_XBULK_PEER=1 BSM_TUNE_WCHUNK=4096
