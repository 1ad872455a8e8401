# feature-env step kernel: occupancy sweep through -DFEAT_MIN_BLOCKS (run on the GPU box)
for mb in 3 4 5 6; do
  python -c "
from contracts_b200 import build
build.build(force=True, extra_flags=['-DFEAT_MIN_BLOCKS=$mb'])" > /dev/null 2>&1 || { echo "min blocks $mb: build failed"; continue; }
  cuobjdump -res-usage contracts_b200/libssd_b200.so 2>/dev/null | grep -A1 "feat_step" | grep -oE "REG:[0-9]+ STACK:[0-9]+"
  python - <<PY 2>&1 | cut -c1-250
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as b
from contracts_b200.features import BatchedFeatureEnv
b.run("cleanup features minblocks $mb", BatchedFeatureEnv("cleanup", 131072, 8, contract="CleanupContract"), 8, 8, 273.125, 200)
b.run("harvest features minblocks $mb", BatchedFeatureEnv("harvest", 131072, 8, contract="HarvestFeaturemodLocalContract"), 8, 7, 321.125, 200)
PY
done
python -c "
from contracts_b200 import build
build.build(force=True)"
