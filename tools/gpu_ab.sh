set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in lib lib_r1 lib_r3; do echo "== $v"; RSSYNC_B200_LIB=$PWD/rs-sync_b200/$v/librssync_b200.so python bench.py --steps 10 --warmup 3 --no-cpu --no-sync 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"; done
echo "== sync C2"; python tools/prof_sync.py C2 2>&1 | tail -2
echo "== sync C4"; python tools/prof_sync.py C4 2>&1 | tail -1
