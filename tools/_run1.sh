set -x
md5sum open3d_slam_b200/*.so > gpurun_out/r2l_md5.txt
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r2l_tests.log
python bench.py --no-extras --no-cpu-baseline --sweep 1 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
