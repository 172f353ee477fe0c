#!/bin/bash
# one iteration on the GPU box: parity tests, secondary configs, the C2 bench and a C4 launch list
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for wl in C4 C5sparse C3 C5dense; do
  timeout 300 python bench_configs.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_cfg_${wl}.json 2>>gpurun_out/${TAG}_cfg.err
  python - <<PY
import json
for line in open("gpurun_out/${TAG}_cfg_${wl}.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("${wl}", "value %.1f ms/step %.4f pack %.1f (%.3f)" % (d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"]), d["roofline"]["stage_ms"])
        break
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_bench.json').readline())
print('C2 value %.1f e2e %.1f ms/step %.4f pack frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']), d['roofline']['stage_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_C4.csv python bench_configs.py --workload C4 --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_C4.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_C4.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(list)
for r in rows: t[r[4].split("(")[0]].append(float(r[-1].replace(",", "")))
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])): print("%-40s n=%3d  last=%.1f us  sum=%.1f" % (k[:40], len(v), v[-1] / 1000, sum(v) / 1000))
PY
