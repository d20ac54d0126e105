# pipelining experiments of the bench graph: input sets (graph branches) x joined / independent train and eval chains
for st in 20 200; do for ns in 3 4; do for sp in 0 1 0 1; do
echo -n "steps=$st NSETS=$ns SPLIT=$sp  "
KODB_BENCH_NSETS=$ns KODB_BENCH_SPLIT=$sp python bench.py --steps $st --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['one_stream_ms_per_step'])"
done; done; done
