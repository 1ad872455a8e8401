# A/B: observe kernel register cap x step chunking (short benches, no CPU arm)
tag=${1:-r2s}
mkdir -p gpurun_out
for v in r80 r72; do
  for c in 1 2 4 8; do
    SSD_LIB_PATH=$PWD/build_variants/libssd_$v.so SSD_STEP_CHUNKS=$c timeout 300 python bench.py --steps 300 --warmup 50 --no-cpu --e2e-steps 100 > gpurun_out/${tag}_${v}_c$c.json 2> gpurun_out/${tag}_${v}_c$c.err
  done
done
SSD_LIB_PATH=$PWD/build_variants/libssd_r72.so timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
