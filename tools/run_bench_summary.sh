python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['e2e']['ms_per_step'], d['e2e'].get('pcie_ceiling_ms'))
for c in d['configs']: print(c['name'][:60], c.get('kernel_ms'), c.get('frac'))
print(d['checks'])
PY
