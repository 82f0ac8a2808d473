set -x
md5sum open3d_slam_b200/*.so > gpurun_out/r2p_md5.txt
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2p_tests.log
for v in 1 0 1 0; do
B2S_PDL=$v python bench.py --no-extras --no-cpu-baseline --sweep 1 > gpurun_out/r2p_bench_pdl$v.json 2> gpurun_out/r2p_bench_pdl$v.err
cp gpurun_out/r2p_bench_pdl$v.json gpurun_out/r2p_bench_pdl${v}_$RANDOM.json
done
