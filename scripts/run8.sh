run() { # n mode mc flags publish tag
  n=$1; mode=$2; mc=$3; fl=$4; pub=$5; tag=$6
  B200GCN_PUBLISH=$pub B200GCN_FLAGS=$fl B200GCN_EXCHANGE=$mode B200GCN_MULTICAST=$mc timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/b8_$tag.json 2> gpurun_out/b8_$tag.err
  grep "^{" gpurun_out/b8_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$tag', round(d['ms_per_step'],3), round(d['value']/1e9,1), d['config']['parallelism'], d['roofline']['launch_ms_by_position_in_step_rank0'])
"
}
run 8 fused 0 0 kernel n8_fused_bulk
run 8 fused 0 0x1000000 kernel n8_fused_lane
run 8 fused 1 0 kernel n8_fused_mc
run 8 fused-split 0 0 kernel n8_split_bulk
run 8 fused-split 1 0 kernel n8_split_mc
run 8 fused 0 0 copy n8_fused_bulk_cepub
run 8 fused 1 0 copy n8_fused_mc_cepub
