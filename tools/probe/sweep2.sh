# steady-state throughput vs slots x frames per slot
for cfg in "16 4" "24 4" "30 4" "40 4" "30 2" "30 3"; do
  set -- $cfg
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --slots $1 --group $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT slots=$1 group=$2', round(d['value'],1), round(d['e2e']['value'],1))"
done
