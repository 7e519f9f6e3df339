# usage: bash scripts/gpu_prof_refresh.sh <tag>  - reduced profiling pass after a change that leaves k_trace alone: ncu launch list of the
# profiling command and one --set full capture of k_pt_shade (three launches); the command runs once without ncu first
TAG=${1:-r02b}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 200 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "profiling command failed without ncu"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled -k regex:k_pt_shade -s 18 -c 3 -o gpurun_out/${TAG}_shade $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out | grep ${TAG}
