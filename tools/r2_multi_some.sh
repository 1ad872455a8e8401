# multi-GPU bench lines of the configs named after N and the tag
N=${1:-2}; tag=${2:-r2ms}; shift 2
mkdir -p gpurun_out
for c in "$@"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 300 --warmup 50 --no-cpu > gpurun_out/${tag}_n${N}_$c.json 2> gpurun_out/${tag}_n${N}_$c.err
done
