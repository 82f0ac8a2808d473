set -x
md5sum open3d_slam_b200/*.so > gpurun_out/r2h_md5.txt
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r2h_tests.log
python bench.py --sweep 1,4,8,32 > gpurun_out/r2h_bench_full.json 2> gpurun_out/r2h_bench_full.err
