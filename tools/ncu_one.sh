# usage: bash tools/ncu_one.sh <kernel-regex> <ncu_targets.py args...>   -> gpurun_out/ncu_<regex>_raw.csv + _sass.csv.gz
set -x
K=$1; shift
mkdir -p gpurun_out
ncu -f --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/one python tools/ncu_targets.py "$@" > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/ncu_one.log
ncu -i /tmp/one.ncu-rep --page raw --csv > gpurun_out/ncu_${K}_raw.csv
ncu -i /tmp/one.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${K}_sass.csv.gz
