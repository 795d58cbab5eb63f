mkdir -p gpurun_out
for s in 3 4 6 8; do BENCH_STREAMS=$s timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', $s, 'value', j['value'], 'e2e', j['e2e']['value'])"; done | tee gpurun_out/streams.log
