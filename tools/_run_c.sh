set -x
python -m pytest tests/test_gpu_parity.py -q -m gpu -k 'voxel or process_scan or config3 or golden or submap_insert' 2>&1 | tail -3 > gpurun_out/r02_voxel_tests.log
python tools/config3_microbench.py --out gpurun_out/r02_config3.json > /dev/null 2>&1
BARGS="--chains 1 --no-graph --steps 2 --warmup 1 --scans-per-step 5 --no-extras --no-sweep --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'icp_kernel|normals_select2' -s 252 -c 2 -o gpurun_out/r02_top2_src -f python bench.py $BARGS > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:'radix|cluster_sort|voxel_mean|voxel_keys|normals_select2|normals_phase2|normals_finish|grid_scatter|grid_count|scan_lookback|seg_head' -c 30 -o gpurun_out/r02_config3_full -f python tools/config3_microbench.py --reps 1 --out "" > gpurun_out/r02_ncu5.log 2>&1
ncu -i gpurun_out/r02_config3_full.ncu-rep --page raw --csv > gpurun_out/r02_config3_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'carve|dense' -s 8 -c 14 -o gpurun_out/r02_carve_full -f python tools/config5_mapper.py --scans 60 > gpurun_out/r02_ncu6.log 2>&1
ncu -i gpurun_out/r02_carve_full.ncu-rep --page raw --csv > gpurun_out/r02_carve_full_raw.csv 2>/dev/null
ls -la gpurun_out
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 58 ]; then rm -f gpurun_out/r02_config3_full.ncu-rep; fi
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 58 ]; then rm -f gpurun_out/r02_carve_full.ncu-rep; fi
du -sh gpurun_out
