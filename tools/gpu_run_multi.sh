# Multi-GPU pass (gpurun --gpus 8): the multi-device driver tests, the copy ceiling and the bench line at N = 1, 2, 4, 8.
# usage: gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_run_multi.sh <tag>'
set -x
cd $GRAFT_REPO_ROOT
TAG=$1
nvidia-smi -L | head -8; nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -q -k "multi_device or whole_file or cli" > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_multi.log
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
  timeout 300 $L tools/pcie_probe_multi.py > gpurun_out/${TAG}_pcie_n$N.json 2> gpurun_out/${TAG}_pcie_n$N.err; tail -c 400 gpurun_out/${TAG}_pcie_n$N.json
  timeout 600 $L bench.py --gpus $N --steps 20 --warmup 3 --no-e2e-file --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -c 700 gpurun_out/${TAG}_bench_n$N.json
done
