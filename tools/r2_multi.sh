# multi-GPU bench lines (one process per GPU via torchrun), every config; N = $1
N=${1:-2}
tag=${2:-r2m}
mkdir -p gpurun_out
for c in cleanup8 features1m harvestfeat1m selfdrive8 harvest16k; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 300 --warmup 50 > gpurun_out/${tag}_n${N}_$c.json 2> gpurun_out/${tag}_n${N}_$c.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_n${N}_ref.json 2> gpurun_out/${tag}_n${N}_ref.err
nvidia-smi topo -m > gpurun_out/${tag}_n${N}_topo.txt 2>&1
