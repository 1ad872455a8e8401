for cfg in "32 3" "32 8" "16 4" "16 13" "8 6" "8 9" "24 5"; do set -- $cfg; SSD_G32_CHUNK=$1 SSD_G32_WARPS=$2 timeout 200 python bench.py --steps 200 --warmup 300 --no-cpu --e2e-steps 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $1 warps $2', '%.3e'%d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])" ; done
