# where the observe kernel's time goes: rebuild with one phase compiled out (results are wrong, timing only)
for flag in ${FLAGS:-NONE OBS_SKIP_SPAWN OBS_SKIP_TRANSPOSE OBS_SKIP_GATHER}; do
  python -c "
from contracts_b200 import build
build.build(force=True, extra_flags=['-D$flag'])" > /dev/null 2>&1 || { echo "$flag: build failed"; continue; }
  timeout 200 python bench.py --steps 200 --warmup 300 --no-cpu --e2e-steps 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$flag', d['roofline']['kernel_ms'], [round(k['ms'],4) for k in d['roofline']['kernels']])"
done
python -c "
from contracts_b200 import build
build.build(force=True)"
