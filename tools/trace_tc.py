"""Per-role clock64() timeline of conv_tc_kernel (CTA 0, pipeline 0, first 48 tiles).

The probes are compiled in only with -DSB200_TC_TRACE_BUILD, so build a second library and point the loader at it:

  SB200_LIB_OUT=$PWD/sonata_b200/lib/libsonata_b200_trace.so SB200_OBJ_DIR=obj_trace \
      SB200_NVCC_EXTRA=-DSB200_TC_TRACE_BUILD python sonata_b200/build.py
  SB200_LIB=$PWD/sonata_b200/lib/libsonata_b200_trace.so python tools/trace_tc.py [cin k dil]

SB200_DEBUG_RES_IS_X=1 aliases the residual to the conv input (the ResBlock situation; timing only).  Columns:
producer issue / landed-as-seen-by-the-producer / converted | MMA a_full seen / issued | epilogue acc_full seen /
TMEM read / done.  A consumer-side "landed" stamp says when the consumer LOOKED, not when the data arrived.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB200_TC_TRACE"] = "1"
from conv_unit import run_case
c = int(sys.argv[1]) if len(sys.argv) > 1 else 32
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rows = 148 * 128 * 100
e, msg = run_case(1, rows, c, c, k, d, 0.1, 0, True, 1.0, False, None)
print("err", e, msg)
