for ns in 2 3 4 6; do for sp in 0 1; do
echo "NSETS=$ns SPLIT=$sp"
KODB_BENCH_NSETS=$ns KODB_BENCH_SPLIT=$sp python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['one_stream_ms_per_step'])"
done; done
