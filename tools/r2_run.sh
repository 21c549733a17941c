set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -12
bash tools/sanitize.sh 2>&1 | tail -30
