#!/bin/bash
# tuning aid: parity of the neighbour paths on the default build, then bench.py --quick per-stage times for it and every libpsolver_<variant>.so
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
bash scripts/bench_variants.sh 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    st = d.get('stage_ms_per_launch', {})
    print(d.get('lib', '').split('/')[-1], 'ms/step', round(d['ms_per_step'], 3), {n: v for n, v in st.items() if n in ('lambda', 'delta_p', 'sort')})
"
