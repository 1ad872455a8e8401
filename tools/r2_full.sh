# full single-GPU validation: tests, smoke, every config, both window sizes, reference arm
tag=${1:-r2full}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench20.json 2> gpurun_out/${tag}_bench20.err
for c in harvest16k features1m harvestfeat1m selfdrive8; do
  timeout 300 python bench.py --config $c --steps 300 --warmup 30 > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
