set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1
ncu --set full --clock-control none -k regex:'rs_hist|rs_scatter|voxel_mean|normals_select2|normals_finish|normals_phase2|grid_scatter|grid_count' -s 11 -c 16 -o gpurun_out/r02_config3_full -f python tools/config3_microbench.py --reps 1 --out "" > gpurun_out/r02_ncu5.log 2>&1
ncu -i gpurun_out/r02_config3_full.ncu-rep --page raw --csv > gpurun_out/r02_config3_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'carve_' -s 40 -c 8 -o gpurun_out/r02_carve_full -f python tools/config5_mapper.py --scans 40 > gpurun_out/r02_ncu6.log 2>&1
ncu -i gpurun_out/r02_carve_full.ncu-rep --page raw --csv > gpurun_out/r02_carve_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'dense_' -s 20 -c 8 -o gpurun_out/r02_dense_full -f python tools/config5_mapper.py --scans 40 > gpurun_out/r02_ncu7.log 2>&1
ncu -i gpurun_out/r02_dense_full.ncu-rep --page raw --csv > gpurun_out/r02_dense_full_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
