#!/bin/bash
# usage: lb_sweep.sh  (runs C2 and a reduced C4 with the library as built)
python bench.py --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], {a:b['ms'] for a,b in d['roofline']['kernels'].items()})"
python bench.py --config C4 --cells 60000 --steps 10 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4/60k', d['ms_per_step'], {a:b['ms'] for a,b in d['roofline']['kernels'].items()})"
