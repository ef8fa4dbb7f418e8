"""Per-role timeline of conv_tc_kernel (CTA 0, pipeline 0): SB200_TC_TRACE=1 python tools/trace_tc.py [cin k dil]"""
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
