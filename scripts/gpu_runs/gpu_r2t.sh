# full GPU test suite (C++ host demo with the fracture hand-off included) + per-kernel times of the particle tick in the e2e loop
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2t_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_gputests.log; tail -8 gpurun_out/r2t_gputests.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:particles -c 300 --csv --log-file gpurun_out/r2t_part_launches.csv python scripts/e2e_stages.py 8192 3 > gpurun_out/r2t_part.log 2>&1; tail -3 gpurun_out/r2t_part.log
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2t_part_launches.csv')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            try: agg[d['Kernel Name'][:50]].append(float(d['Metric Value'].replace(',','')))
            except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(k, len(v), 'total us', round(sum(v)/1000,1), 'max us', round(max(v)/1000,1), 'last7', [round(x/1000,1) for x in v[-7:]])
PY
python bench.py --workload column --size 2048 --steps 1000 --warmup 3 > gpurun_out/r2t_column.json 2> gpurun_out/r2t_column.err; tail -2 gpurun_out/r2t_column.err
python -c "
import json
d=json.loads(open('gpurun_out/r2t_column.json').read().strip().splitlines()[-1]); print('column', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])"
