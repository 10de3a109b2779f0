for g in 8 4 2 0; do
  echo "== align $g"
  NMPC_B200_ALIGN_GROUP=$g python tools/iter_cost_probe.py t_trajectory 16384 | tail -1
  NMPC_B200_ALIGN_GROUP=$g python tools/iter_cost_probe.py nmpc_tt 16384 | tail -1
  NMPC_B200_ALIGN_GROUP=$g python bench.py --config 2 --steps 40 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('cfg2', d['value'], d['ms_per_step'])"
done
