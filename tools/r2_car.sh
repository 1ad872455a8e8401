# selfdrive: parity tests then the selfdrive config (+ optional ncu capture)
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_selfdrive_golden.py tests/test_negotiate_golden.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
timeout 300 python bench.py --config selfdrive8 --steps 300 --warmup 30 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_selfdrive8.json 2> gpurun_out/${tag}_selfdrive8.err
if [ -n "$PROF" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:car_kernel -s 430 -c 1 -f -o gpurun_out/${tag}_car python bench.py --config selfdrive8 --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2 > gpurun_out/${tag}_car.log 2>&1
fi
