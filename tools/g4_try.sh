# Stream-K lone-launch kernel: parity (bounded by timeout), A/B bench lines, steady-state timeline.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for g4 in 1 0; do
  TMAC_B200_G4=$g4 timeout -k 10 200 python bench.py --steps 30 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('G4=$g4 value %.0f GB/s (%.2f us/layer) | lone %.2f us frac %.3f | two-call %.0f GB/s | grouped %.2f us | e2e %.0f | %s' % (d['value'], d['ms_per_step']*1e3/32, r['us_per_launch'], r['frac'], r['two_call_step']['GBps'], r.get('grouped_launch',{}).get('us_per_gemv',0), d['e2e']['value'], r['launch']))
"
done
timeout 120 python tools/trace_steady.py 2>&1 | tail -5; TRACE_FUSED=1 timeout 120 python tools/trace_steady.py 2>&1 | tail -4
