for h in 1024 1536 2048 3072; do
echo -n "HOT_TARGET=$h  "
KODB_HOT_TARGET=$h python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['one_stream_ms_per_step'], d['stages_ms']['cand_floor_kernel'], d['stages_ms']['select_nms_kernel'])"
done
