# The round's standard GPU pass (one gpurun call): smoke, the GPU parity suite, the bench lines of the three single-GPU
# configs + the reference arm, the ncu launch list of the bench command and one full capture of the dominant kernel per
# config.  usage: gpurun --timeout 3000 -- 'bash tools/gpu_full_run.sh <tag> [notests]'
set -x
cd $GRAFT_REPO_ROOT
TAG=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$2" != "notests" ]; then
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_ref.json
timeout 300 python bench.py --workload c1_16bit_44k1_60s --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c1.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_c1.json
timeout 300 python bench.py --workload c3_32bit_192k_600s --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --profile > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG}_c2 python bench.py --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG}_c3 python bench.py --workload c3_32bit_192k_600s --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG}_c1 python bench.py --workload c1_16bit_44k1_60s --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu_c1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zf_encode_stereo_kernel -c 1 -o gpurun_out/prof_${TAG}_tail python bench.py --workload c1_16bit_44k1_60s --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu_tail.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_c1.csv python bench.py --workload c1_16bit_44k1_60s --steps 3 --warmup 3 --profile > gpurun_out/${TAG}_ncu_c1l.log 2>&1
timeout 300 python bench.py --workload c4_16bit_44k1_60s_lpc12 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_c4.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_c4.json
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:zf_dec -c 8 --csv --log-file gpurun_out/${TAG}_launches_decode.csv python tools/decode_probe.py c2 1 > gpurun_out/${TAG}_ncu_decl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zf_dec_frames -c 1 -o gpurun_out/prof_${TAG}_decode python tools/decode_probe.py c2 1 > gpurun_out/${TAG}_ncu_decode.log 2>&1
for c in c1 c2 c3; do timeout 300 python tools/decode_probe.py $c 3 2>&1 | tail -2 | head -1; done > gpurun_out/${TAG}_decode_probe.txt; cat gpurun_out/${TAG}_decode_probe.txt
tail -2 gpurun_out/${TAG}_ncu_c3.log
