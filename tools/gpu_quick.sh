# quick GPU iteration: parity tests, then the C2 grid kernel time for each library given
# usage: bash tools/gpu_quick.sh [tests|notests] [lib.so ...]
mode=${1:-tests}; shift
mkdir -p gpurun_out
if [ "$mode" = tests ]; then python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -4; fi
if [ $# -eq 0 ]; then set -- rs-sync_b200/lib/librssync_b200.so; fi
for lib in "$@"; do
  echo "== $lib"
  RSSYNC_B200_LIB=$PWD/$lib python tools/prof_presync.py C2 4 2>&1 | tail -3
done
