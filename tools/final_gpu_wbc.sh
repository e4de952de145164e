set -x
mkdir -p gpurun_out/r2q
ncu --set full --clock-control none --kernel-name regex:^k_ --launch-skip 8 --launch-count 8 -f -o gpurun_out/r2q/prof_wbc python tools/wbc_throughput.py > gpurun_out/r2q/prof_wbc.log 2>&1
ls -la gpurun_out/r2q
