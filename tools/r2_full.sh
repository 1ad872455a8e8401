# full single-GPU validation: tests, smoke, every config, both window sizes, reference arm (every stage under a timeout)
tag=${1:-r2full}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 --timeout 120 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
timeout 150 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench20.json 2> gpurun_out/${tag}_bench20.err
for c in harvest16k features1m harvestfeat1m selfdrive8; do
  timeout 200 python bench.py --config $c --steps 300 --warmup 30 > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
