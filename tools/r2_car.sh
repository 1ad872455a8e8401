# selfdrive: parity tests then the selfdrive config with the in-tree library and the variants named after the tag (+ optional ncu capture)
tag=${1:-r2c}; shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_selfdrive_golden.py tests/test_negotiate_golden.py -m gpu -q -x --timeout 100 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
for v in base "$@"; do
  if [ $v = base ]; then unset SSD_LIB_PATH; else export SSD_LIB_PATH=$PWD/build_variants/libssd_$v.so; fi
  timeout 150 python bench.py --config selfdrive8 --steps 300 --warmup 30 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_${v}_selfdrive8.json 2> gpurun_out/${tag}_${v}_selfdrive8.err
done
unset SSD_LIB_PATH
if [ -n "$PROF" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:car_kernel -s 430 -c 1 -f -o gpurun_out/${tag}_car python bench.py --config selfdrive8 --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2 > gpurun_out/${tag}_car.log 2>&1
fi
