set -x
md5sum open3d_slam_b200/*.so > gpurun_out/r02_final_md5.txt
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python tools/icp_phases.py > gpurun_out/r02_icp_phases.txt 2>&1
B2S_ICP_MAX_CLUSTER=16 python tools/icp_phases.py > gpurun_out/r02_icp_phases_16sm.txt 2>&1
BARGS="--chains 1 --no-graph --steps 2 --warmup 1 --scans-per-step 5 --no-extras --no-sweep --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 5300 -c 420 --csv --log-file gpurun_out/r02_launches.csv python bench.py $BARGS > gpurun_out/r02_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_config3_launches.csv python tools/config3_microbench.py --reps 2 --out "" > gpurun_out/r02_ncu4.log 2>&1
du -sh gpurun_out
