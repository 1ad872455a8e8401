# feature envs: parity tests, the two feature configs with the in-tree library and the variants named in $2.., then an ncu capture
tag=${1:-r2g}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_features_vs_oracle.py tests/test_features_golden.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
for v in base "$@"; do
  for c in features1m harvestfeat1m; do
    if [ $v = base ]; then unset SSD_LIB_PATH; else export SSD_LIB_PATH=$PWD/build_variants/libssd_$v.so; fi
    timeout 300 python bench.py --config $c --steps 300 --warmup 30 --no-cpu --e2e-steps 50 > gpurun_out/${tag}_${v}_$c.json 2> gpurun_out/${tag}_${v}_$c.err
  done
done
unset SSD_LIB_PATH
if [ -n "$PROF" ]; then bash tools/r2_prof_feat.sh ${tag}; fi
