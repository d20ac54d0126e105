#!/bin/bash
# parity of the eval tail + short bench + instruction counters of the candidate kernel: tools/quick_eval.sh <tag>
T=$1
python -m pytest tests/test_eval_tail_gpu.py tests/test_experiment_gpu.py -x -q 2>&1 | tail -2
for v in 0 1 0 1; do
echo -n "KODB_EVAL_TMA=$v "
KODB_EVAL_TMA=$v python bench.py --steps 60 --warmup 5 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'one_stream', d['one_stream_ms_per_step'], 'train/cand/sel', d['stages_ms']['train_tail_kernel'], d['stages_ms']['cand_floor_kernel'], d['stages_ms']['select_nms_kernel'], 'full', d['tiles_listed_in_full'])"
done
tools/ncu_quick.sh cand_floor gpurun_out/q_cand_$T.csv
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_cand_$T.csv')) if len(r)>10]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value')
acc={}
for r in rows[1:]:
    acc.setdefault(r[mi],[]).append(float(r[vi].replace(',','')))
for k,v in acc.items(): print(k, sum(v)/len(v))
PY
