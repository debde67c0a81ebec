for cfg in "30 1" "30 2" "30 3" "30 4" "20 4" "32 4"; do
  set -- $cfg
  VSC_MARCH_SMS=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --slots $1 --group $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT slots=$1 group=$2', round(d['value'],1), round(d['e2e']['value'],1))"
done
