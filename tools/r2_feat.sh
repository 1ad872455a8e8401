tag=${1:-r2feat}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_features_golden.py tests/test_cuda_solver_joint.py -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
for v in feat4 feat5 feat6; do
  for c in features1m harvestfeat1m; do
    SSD_LIB_PATH=$PWD/build_variants/libssd_$v.so timeout 300 python bench.py --config $c --steps 300 --warmup 30 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_${v}_$c.json 2> gpurun_out/${tag}_${v}_$c.err
  done
done
