# Which knob removes the data-parallel overhead at N GPUs?  bash tools/dp_overhead_probe.sh N
N=$1
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-tf32 --no-extras --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'])"
}
run baseline A=1
run nch2 NCCL_MAX_NCHANNELS=2
run nch2_res2 NCCL_MAX_NCHANNELS=2 NANOGRAD_B200_SM_RESERVE=2
run nch4_res4 NCCL_MAX_NCHANNELS=4 NANOGRAD_B200_SM_RESERVE=4
run nch1_res1 NCCL_MAX_NCHANNELS=1 NANOGRAD_B200_SM_RESERVE=1
run nch2_res2_chunk18 NCCL_MAX_NCHANNELS=2 NANOGRAD_B200_SM_RESERVE=2 NANOGRAD_B200_DP_CHUNK_FLOATS=262144
run res8 NANOGRAD_B200_SM_RESERVE=8
