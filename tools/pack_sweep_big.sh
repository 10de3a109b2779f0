# configs 3-5: does splitting a big batch into pipelined sub-batches help the device arm too?
for c in 3 4 5; do for S in 1 2 4 8; do
  python bench.py --config $c --pipelines $S --no-cpu-baseline --no-free-running --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('cfg $c S=$S', round(d['value']), round(d['ms_per_step'],1), 'p50', round(d['p50_step_ms'],1))"
done; done
