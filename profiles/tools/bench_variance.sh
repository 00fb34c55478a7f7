# value / with_metrics / e2e of the quick bench, a few times, with the NVML clock sampler at its default period and slowed down
for period in 0.05 10 0.05 10; do
  BG_CLOCK_PERIOD=$period python bench.py --no-cpu-baseline --no-extra-workloads --no-hbm-roofline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('period $period', d['value'], d['with_metrics']['value'], d['e2e']['value'], d['clocks']['samples'])"
done
