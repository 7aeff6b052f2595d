# r02 GPU call A: full test suite (incl. full-size parity), smoke, baseline bench, sanitizer passes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -3 gpurun_out/bench_r02_a.err; cut -c1-600 gpurun_out/bench_r02_a.json
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python tools/sanitize_workload.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool exit $?"; tail -2 gpurun_out/sanitizer_$tool.out; tail -4 gpurun_out/sanitizer_$tool.log
done
