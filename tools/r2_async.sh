# async host paths of the feature / selfdrive envs: tests + the three configs
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_features_vs_oracle.py tests/test_selfdrive_golden.py tests/test_cuda_properties.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${tag}_tests.log
for c in features1m harvestfeat1m selfdrive8; do
  timeout 300 python bench.py --config $c --steps 300 --warmup 30 --no-cpu --e2e-steps 300 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err
done
