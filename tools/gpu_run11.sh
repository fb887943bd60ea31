set -x
cd $GRAFT_REPO_ROOT
for v in c3 c4 c3 c4; do
cp zig-flac_b200/lib_$v.so.bin zig-flac_b200/libzigflac_b200.so
timeout 300 python bench.py --workload x1_16bit_44k1_3600s --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('16-bit $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity'])"
done
