"""Per-role timeline (tools/trace_tc.py) for an arbitrary conv shape:
  SB200_LIB=.../libsonata_b200_trace.so python tools/trace_tc_shape.py rows cin cout k dil act res"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB200_TC_TRACE"] = "1"
from conv_unit import run_case
rows, cin, cout, k, d, act, res = (int(x) for x in sys.argv[1:8])
e, msg = run_case(1, rows, cin, cout, k, d, 1.0, act, bool(res), 1.0, False, None)
print("err", e, msg)
