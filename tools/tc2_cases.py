import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tc2_check as t
for (n, Q) in [(200000, 2500), (200000, 2500), (300000, 1024), (300000, 1024), (1000000, 1024), (1000000, 1024)]:
    t.run(20, 60, n, Q, 10)
