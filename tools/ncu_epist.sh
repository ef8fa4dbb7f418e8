set -x
for v in epist noepist; do
  if [ $v = noepist ]; then export SB200_TC_NOEPIST=1; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 235 -c 4 -f -o gpurun_out/prof_r02_$v python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c5 --no-secondary > gpurun_out/ncu_r02_$v.log 2>&1
  ncu -i gpurun_out/prof_r02_$v.ncu-rep --page raw --csv > gpurun_out/prof_r02_${v}_raw.csv 2>/dev/null
done
ls -la gpurun_out/prof_r02_*
