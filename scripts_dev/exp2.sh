for ch in 1 2 3; do for a in admm fista; do
LPC_CHAINS=$ch timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --algo $a 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chains=$ch $a', round(d['value'],2), round(d['ms_per_step'],1), round(d['e2e']['value'],2))"
done; done
