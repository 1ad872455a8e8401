# gridworld step: full gpu test-suite, then cleanup8 (per-kernel times) and harvest16k
tag=${1:-r2gr}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config harvest16k --steps 300 --warmup 30 --no-cpu > gpurun_out/${tag}_harvest16k.json 2> gpurun_out/${tag}_harvest16k.err
