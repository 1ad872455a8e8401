tag=${1:-r2t3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config harvest16k --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench_harvest16k.json 2> gpurun_out/${tag}_bench_harvest16k.err
SSD_OBS_GENERIC=1 SSD_LOGIC_GENERIC=1 timeout 300 python bench.py --config harvest16k --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench_harvest16k_generic.json 2> gpurun_out/${tag}_bench_harvest16k_generic.err
