python -c "import torch; print(torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream,'priority_range') else 'n/a')"
for sp in 0 -1 -2 -5 0 -5; do
echo -n "SEL_PRIO=$sp  "
KODB_BENCH_SEL_PRIO=$sp python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['map_report'])"
done
