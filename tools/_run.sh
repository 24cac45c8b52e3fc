timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py > gpurun_out/bench17.json 2> gpurun_out/bench17.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench17.json'))
print(d['value'], d['e2e']['value'], d['dispatch'])
PY
