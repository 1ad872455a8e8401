# last check of the committed state: gpu test-suite + smoke
tag=${1:-r2last}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x --timeout 100 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
