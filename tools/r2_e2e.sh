# host-path tests + e2e of a small-batch and a large-batch config
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_properties.py tests/test_features_vs_oracle.py tests/test_selfdrive_golden.py tests/test_vector_env.py -m gpu -q -x --timeout 100 -k "host or async or prefix or vector" 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
for c in harvest16k cleanup8; do
  timeout 200 python bench.py --config $c --steps 300 --warmup 30 --no-cpu --e2e-steps 300 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err
done
