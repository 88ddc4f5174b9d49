for i in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', round(d['value'],1), d['ms_per_step'], d['step_ms'])"
python bench.py --steps 10 --warmup 3 --queries-per-gpu 16 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('Q16', round(d['value'],1), d['ms_per_step'], d['step_ms'])"
done
