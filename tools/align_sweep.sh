# alignment study: NMPC_B200_ALIGN_GROUP (warps per alignment group) and NMPC_B200_ALIGN_MID (mid-iteration points, bit mask)
for m in 1 5 7; do
  echo "== align mid $m"
  NMPC_B200_ALIGN_MID=$m python tools/iter_cost_probe.py t_trajectory 16384 | tail -1
  NMPC_B200_ALIGN_MID=$m python tools/iter_cost_probe.py nmpc_tt 16384 | tail -1
  NMPC_B200_ALIGN_MID=$m python tools/iter_cost_probe.py race_track_2 16384 30 | tail -1
  NMPC_B200_ALIGN_MID=$m python bench.py --config 2 --steps 40 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('cfg2', d['value'], d['ms_per_step'])"
done
