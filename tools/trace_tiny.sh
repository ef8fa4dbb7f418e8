# launch-level and per-stage stamps of single-tile launches (rows 40..46: stage j; row 47: kernel entry / prologue done /
# dependency wait done / before TMEM dealloc), see tools/trace_tc.py for the build
export SB200_LIB=$PWD/sonata_b200/lib/libsonata_b200_trace.so
for cfg in "1100 192 384 5 1 2 0" "1100 192 384 1 1 0 0"; do echo "=== $cfg" >> gpurun_out/trace_tiny.txt; timeout 100 python tools/trace_tc_shape.py $cfg 2>&1 | grep -v ": *-1 *-1 *-1 *-1 *-1 *-1 *-1 *-1" >> gpurun_out/trace_tiny.txt; done
