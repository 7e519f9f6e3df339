# usage: bash scripts/gpu_bench_prof.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --profile"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 2 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pt_shade -s 18 -c 3 -f -o gpurun_out/${TAG}_shade $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
tail -3 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/
