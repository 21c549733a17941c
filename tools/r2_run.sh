set -x
mkdir -p gpurun_out
bash tools/sanitize.sh 2>&1 | tail -24
