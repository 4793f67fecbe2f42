set -x
mkdir -p gpurun_out
python tests/ref_gpu_baseline.py > gpurun_out/ref_gpu_baseline.json 2> gpurun_out/ref_gpu_baseline.err
tail -3 gpurun_out/ref_gpu_baseline.err
ncu --set full --clock-control none --import-source on -k regex:'rotate_kernel|gauss_stream|bilinear2|generic_scale|median_net|rgb2yuv_kernel|flip' -o /tmp/r2 python tools/ncu_targets.py bilinear generic rotate gauss median rgb2yuv fliph > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
ncu -i /tmp/r2.ncu-rep --page raw --csv > gpurun_out/r1b_ncu_full_raw.csv
for k in rotate_kernel gauss_stream bilinear2 generic_scale; do ncu -i /tmp/r2.ncu-rep -k regex:$k --page source --csv 2>/dev/null | gzip > gpurun_out/r1b_${k}_sass.csv.gz; done
ls -la gpurun_out
