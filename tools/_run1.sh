set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_nn.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c_tests.log
python tools/icp_phases.py > gpurun_out/r2c_phases8.log 2>&1
B2S_ICP_MAX_CLUSTER=16 python tools/icp_phases.py > gpurun_out/r2c_phases16.log 2>&1
B2S_LIB=open3d_slam_b200/libb2s_alt512.so B2S_ICP_MAX_CLUSTER=16 python tools/icp_phases.py > gpurun_out/r2c_phases16_alt512.log 2>&1
B2S_LIB=open3d_slam_b200/libb2s_alt768.so B2S_ICP_MAX_CLUSTER=16 python tools/icp_phases.py > gpurun_out/r2c_phases16_alt768.log 2>&1
python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench8.json 2> gpurun_out/r2c_bench8.err
B2S_ICP_MAX_CLUSTER=16 python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench16.json 2> gpurun_out/r2c_bench16.err
B2S_LIB=open3d_slam_b200/libb2s_alt512.so B2S_ICP_MAX_CLUSTER=16 python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench16_alt512.json 2> gpurun_out/r2c_bench16_alt512.err
B2S_LIB=open3d_slam_b200/libb2s_alt768.so B2S_ICP_MAX_CLUSTER=16 python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench16_alt768.json 2> gpurun_out/r2c_bench16_alt768.err
B2S_NORMALS_CELL_FACTOR=3 python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench8_cf3.json 2> gpurun_out/r2c_bench8_cf3.err
B2S_NORMALS_CELL_FACTOR=2.5 python bench.py --no-extras --no-cpu-baseline --sweep 1,16 > gpurun_out/r2c_bench8_cf25.json 2> gpurun_out/r2c_bench8_cf25.err
