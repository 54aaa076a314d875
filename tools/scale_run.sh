# Scaling runs on one box with N GPUs visible: bash tools/scale_run.sh N   (N = 2, 4 or 8)
N=$1
set -x
python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2_dist_pytest_n$N.log 2>&1
tail -3 gpurun_out/r2_dist_pytest_n$N.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_ref_n1.json 2> gpurun_out/r2_ref_n1.err
for mode in "weak" "weak --sync-bn" "strong" "strong --sync-bn"; do
  tag=$(echo $mode | tr -d ' -')
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --scaling $mode --no-tf32 --no-extras > gpurun_out/r2_scale_n${N}_$tag.json 2> gpurun_out/r2_scale_n${N}_$tag.err
  tail -1 gpurun_out/r2_scale_n${N}_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['scaling'], d['config']['batchnorm'][:24], d['value'], d['ms_per_step'], d['e2e']['value'], d['replica_checksum']['identical_across_ranks'] if d['replica_checksum'] else None)"
done
