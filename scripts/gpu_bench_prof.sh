# usage: bash scripts/gpu_bench_prof.sh <tag>   (run under gpurun, ONE GPU; writes gpurun_out/<tag>_*)
# tests -> bench (both arms) -> ncu launch list of the profiling command -> ncu --set full of the dominant kernels (one launch each), each
# ncu pass only after the same command has exited 0 without ncu
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json; echo; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; head -c 300 gpurun_out/${TAG}_bench_ref.json; echo
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "profiling command failed without ncu"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
NCU="timeout 600 ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
$NCU -k 'regex:k_trace<\(bool\)0, \(int\)0, \(int\)1>' -s 3 -c 1 -o gpurun_out/${TAG}_prof_primary $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
$NCU -k 'regex:k_trace<\(bool\)1, \(int\)0, \(int\)2>' -s 3 -c 1 -o gpurun_out/${TAG}_prof_shadow  $CMD >> gpurun_out/${TAG}_ncu2.log 2>&1
$NCU -k 'regex:k_trace<\(bool\)0, \(int\)0, \(int\)0>' -s 3 -c 1 -o gpurun_out/${TAG}_prof_incoh   $CMD >> gpurun_out/${TAG}_ncu2.log 2>&1
$NCU -k regex:k_pt_shade -s 18 -c 3 -o gpurun_out/${TAG}_shade $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
$NCU -k regex:k_stream_read -s 2 -c 7 -o gpurun_out/${TAG}_l2 $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_ncu2.log
ls gpurun_out/ | grep ${TAG}
