# gpu test-suite + smoke + short bench (one GPU)
tag=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -80 > gpurun_out/${tag}_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
