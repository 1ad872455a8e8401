# one test file / expression under a short timeout
tag=${1:-r2o}; shift
mkdir -p gpurun_out
timeout 400 python -m pytest "$@" -m gpu -q -x --timeout 120 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
