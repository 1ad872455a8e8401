tag=${1:-r2s3}
mkdir -p gpurun_out
for v in w8r80 w7r88 w6r104 w12r80 w10r96 w8r88 w8r96; do
  SSD_LIB_PATH=$PWD/build_variants/libssd_$v.so SSD_DEBUG=1 timeout 300 python bench.py --steps 300 --warmup 50 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
done
