set -x
md5sum open3d_slam_b200/*.so > gpurun_out/r2e_md5.txt
python tools/icp_phases.py > gpurun_out/r2e_phases8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'icp_kernel' -s 12 -c 2 -o gpurun_out/r02c_icp -f python bench.py --chains 1 --no-graph --steps 2 --warmup 2 --scans-per-step 6 --no-extras --no-sweep --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
python bench.py --sweep 1,4,8,32 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err
