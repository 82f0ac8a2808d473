set -x
for g in 2368 592 1184 296 592 2368; do
B2S_NS2_GRID=$g python bench.py --no-extras --no-cpu-baseline --sweep 1 > gpurun_out/r2m_bench_$g.json 2> gpurun_out/r2m_bench_$g.err
cp gpurun_out/r2m_bench_$g.json gpurun_out/r2m_bench_${g}_$RANDOM.json
done
