# guarded validation: every stage under a short timeout, later stages only if the earlier ones passed
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1 || { echo "smoke failed/hung rc=$?" >> gpurun_out/${tag}_smoke.log; exit 0; }
timeout 500 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
grep -q "failed\|error\|Timeout" gpurun_out/${tag}_tests.log && exit 0
timeout 200 python bench.py --steps 500 --warmup 50 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 120 python bench.py --config harvest16k --steps 300 --warmup 30 > gpurun_out/${tag}_harvest16k.json 2> gpurun_out/${tag}_harvest16k.err
