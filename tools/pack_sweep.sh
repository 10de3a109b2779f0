# config 2: sub-batch count x instances per resident warp (how the 4096 closed loops are packed onto the SMs)
for S in 4 8 16 32; do for F in 1 2 4; do
  NMPC_B200_FILL=$F python bench.py --config 2 --steps 60 --warmup 5 --pipelines $S --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('S=$S fill=$F', round(d['value']), round(d['ms_per_step'],2), 'p50', d.get('p50_batch_step_ms'))"
done; done
