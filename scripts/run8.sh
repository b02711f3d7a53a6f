# 8-GPU variant comparison on config 2 (run under `gpurun --gpus 8`): the exchange modes of sharded.ShardedPropagator.
# Round 1 also compared TMA bulk peer stores and a copy-engine ego publish (both slower; removed).
run() { # n mode multicast merge tag
  n=$1; mode=$2; mc=$3; merge=$4; tag=$5
  B200GCN_CHAIN_MERGE=$merge B200GCN_EXCHANGE=$mode B200GCN_MULTICAST=$mc timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/b8_$tag.json 2> gpurun_out/b8_$tag.err
  grep "^{" gpurun_out/b8_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$tag', round(d['ms_per_step'],3), round(d['value']/1e9,1), d['config']['parallelism'], d['roofline'].get('phase_end_us_rank0'), d.get('parity_max_scaled'))
"
}
run 8 chain 1 1 n8_chain_merged
run 8 chain 1 0 n8_chain
run 8 chain 0 1 n8_chain_peer_stores
run 8 fused 1 0 n8_fused_mc
run 8 allgather 0 0 n8_allgather
