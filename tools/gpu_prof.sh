set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
