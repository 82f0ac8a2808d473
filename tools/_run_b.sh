set -x
BARGS="--chains 1 --no-graph --steps 2 --warmup 1 --scans-per-step 5 --no-extras --no-sweep --no-cpu-baseline"
ncu --set full --clock-control none -s 5300 -c 42 -o gpurun_out/r02_scan_full -f python bench.py $BARGS > gpurun_out/r02_ncu2.log 2>&1
ncu -i gpurun_out/r02_scan_full.ncu-rep --page raw --csv > gpurun_out/r02_scan_full_raw.csv 2>/dev/null
ls -la gpurun_out
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 58 ]; then rm -f gpurun_out/r02_scan_full.ncu-rep; fi
du -sh gpurun_out
