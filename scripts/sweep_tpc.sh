#!/bin/bash
for t in "$@"; do
  UNIIR_SEARCH_TPC=$t timeout 200 python bench.py --workload search --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/s.json
  python -c "import json; d=json.load(open('/tmp/s.json')); print('tpc', $t, round(d['value']), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
done
