#!/usr/bin/env python
"""Per-kernel share table from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import launch_shares
print("| share | launches | avg us | kernel |\n|---:|---:|---:|---|")
for k, sh, n, avg in launch_shares(sys.argv[1]):
    print(f"| {sh:.2f}% | {n} | {avg:.2f} | `{k}` |")
