mkdir -p gpurun_out/r2b
for c in 2 3 4 5; do
  python bench.py --config $c > gpurun_out/r2b/bench_c${c}_1gpu.json 2> gpurun_out/r2b/bench_c${c}_1gpu.err; echo "cfg $c rc $?"; tail -c 400 gpurun_out/r2b/bench_c${c}_1gpu.json | head -c 300; echo
done
python bench.py --impl reference --config 2 > gpurun_out/r2b/ref_c2.json 2>/dev/null; echo ref rc $?
