# The commands behind profiles/r02_* and profiles/bench_r02_* (run under gpurun on one B200; the WBC --set full capture is its own
# call, tools/final_gpu_wbc.sh, because of the size limit of what a call may bring back). MPC=1 also re-captures the MPC profiles.
set -x
mkdir -p gpurun_out/r2q
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2q/pytest_gpu.txt; cat gpurun_out/r2q/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/r2q/smoke.txt; cat gpurun_out/r2q/smoke.txt
python bench.py 2>gpurun_out/r2q/bench.err | tail -1 > gpurun_out/r2q/bench_1gpu.json; cut -c1-300 gpurun_out/r2q/bench_1gpu.json
python bench.py --impl reference 2>gpurun_out/r2q/bench_ref.err | tail -1 > gpurun_out/r2q/bench_ref.json; cut -c1-300 gpurun_out/r2q/bench_ref.json
if [ "${MPC:-0}" = "1" ]; then
QMB200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:^k_ -c 400 --csv --log-file gpurun_out/r2q/launches.csv python bench.py --steps 8 --warmup 1 --no-wbc --no-cpu-baseline --no-latency > gpurun_out/r2q/launch_run.log 2>&1
QMB200_GRAPH=0 ncu --set full --import-source on --clock-control none --kernel-name regex:^k_ --launch-skip 33 --launch-count 11 -f -o gpurun_out/r2q/prof python bench.py --steps 2 --warmup 3 --no-wbc --no-cpu-baseline --no-latency > gpurun_out/r2q/prof.log 2>&1
for t in memcheck synccheck racecheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_case.py > gpurun_out/r2q/$t.log 2>&1; tail -1 gpurun_out/r2q/$t.log; done
fi
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:^k_ --csv --log-file gpurun_out/r2q/wbc_launches.csv python tools/wbc_throughput.py > gpurun_out/r2q/wbc_launch_run.log 2>&1
ls -la gpurun_out/r2q
