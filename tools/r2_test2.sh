tag=${1:-r2t2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
timeout 600 python tools/bench_dict_api.py > gpurun_out/${tag}_dict_api.json 2> gpurun_out/${tag}_dict_api.err
for c in harvest16k features1m selfdrive8; do
  timeout 300 python bench.py --config $c --steps 200 --warmup 20 --no-cpu > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
