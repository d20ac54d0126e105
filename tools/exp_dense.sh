for ps in 1 2 3 4 6; do
echo "DENSE_PER_SM=$ps"
KODB_TRAIN_DENSE_PER_SM=$ps python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['one_stream_ms_per_step'], d['stages_ms']['train_tail_kernel'])"
done
