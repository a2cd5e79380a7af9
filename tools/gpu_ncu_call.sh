# usage: tools/gpu_ncu_call.sh <tag> <kernel-regex> <log2n> <kind> [launch-skip] [launch-count]
# one `ncu --set full` capture of the named kernel(s) on one pass of the hot path; report lands in gpurun_out/
mkdir -p gpurun_out
TAG=$1; KRE=$2; LOG2N=${3:-24}; KIND=${4:-gasdark}; SKIP=${5:-0}; CNT=${6:-1}
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"$KRE" --launch-skip $SKIP --launch-count $CNT \
  -f -o gpurun_out/$TAG python tools/ncu_driver.py $LOG2N $KIND > gpurun_out/$TAG.log 2>&1; echo "ncu $TAG rc=$?"
tail -2 gpurun_out/$TAG.log; ls -la gpurun_out/$TAG.ncu-rep
