#!/bin/bash
# ncu captures of round 2 (run under gpurun on ONE GPU; every ncu run follows a plain run of the same command)
set -u
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-also --steps 3 --warmup 3"
run_full() {   # name, kernel regex, bench args...
    local name=$1 rx=$2; shift 2
    $B "$@" > $O/${name}_plain.json 2> $O/${name}_plain.err || { echo "$name: plain run failed"; return; }
    ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o $O/$name $B "$@" > $O/${name}_ncu.log 2>&1
    echo "$name: ncu rc=$?"
}
run_list() {   # name, bench args...
    local name=$1; shift
    $B "$@" > $O/${name}_plain.json 2> $O/${name}_plain.err || { echo "$name: plain run failed"; return; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${name}.csv $B "$@" > $O/${name}_ncu.log 2>&1
    echo "$name: ncu rc=$?"
}
run_full r02_c5_spmm_tma_full spmm_tma --workload c5
run_full r02_c3T_stream_warp_full stream_warp --workload c3 --op T
run_full r02_c3N_stream_warp_full stream_warp --workload c3
run_full r02_c1_stream_warp_full stream_warp --workload c1
run_full r02_c5_nrhs8_spmm_tma_full spmm_tma --workload c5 --nrhs 8
run_list r02_c2_launches --workload c2
run_list r02_c5_launches --workload c5
run_list r02_c1_launches --workload c1
ls -la $O/*.ncu-rep
