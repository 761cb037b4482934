set -u
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also --steps 3 --warmup 3"
run_full() {   # name, kernel regex, bench args...
    local name=$1 rx=$2; shift 2
    $B "$@" > $O/${name}_plain.json 2> $O/${name}_plain.err || { echo "$name: plain run failed"; return; }
    ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o $O/$name $B "$@" > $O/${name}_ncu.log 2>&1
    echo "$name: ncu rc=$?"
}
run_full r02b_c3T_stream_warp_full stream_warp --workload c3 --op T
run_full r02b_c4N_fused_tma_full sym_fused_tma --workload c4 --op N
ls -la $O/*.ncu-rep
