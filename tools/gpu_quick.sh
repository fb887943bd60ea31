# Quick GPU check: selected tests and one bench line.  usage: gpurun -- 'bash tools/gpu_quick.sh <tag> "<pytest -k expr>" [bench args]'
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; KEXPR=$2; shift; shift
timeout 1500 python -m pytest tests -m gpu -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
if [ -n "$1" ]; then timeout 600 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; fi
