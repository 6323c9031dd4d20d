for ry in 1 2 4; do for cv in 0 1; do
LPC_RY=$ry LPC_COLS_VARIANT=$cv timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('RY=$ry CV=$cv', round(d['value'],2), {k: round(v*1e3) for k,v in d['roofline']['phases_ms_per_iter'].items()})"
done; done
