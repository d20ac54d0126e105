# pipelining experiments of the bench graph: number of input sets (graph branches), joined / independent chains, priority of
# the select/NMS streams
for st in 20 200; do for cfg in "3 1 -3" "4 1 -3" "5 1 -3" "3 1 0" "3 0 0"; do set -- $cfg
echo -n "steps=$st NSETS=$1 SPLIT=$2 SEL_PRIO=$3  "
KODB_BENCH_NSETS=$1 KODB_BENCH_SPLIT=$2 KODB_BENCH_SEL_PRIO=$3 python bench.py --steps $st --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['one_stream_ms_per_step'])"
done; done
