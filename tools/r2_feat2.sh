# feature envs: parity tests (fixtures + oracle rollouts) then the two feature configs
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_features_vs_oracle.py tests/test_features_golden.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
for c in features1m harvestfeat1m; do
  timeout 300 python bench.py --config $c --steps 300 --warmup 30 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_$c.json 2> gpurun_out/${tag}_$c.err
done
