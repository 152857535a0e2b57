#!/usr/bin/env python
"""Short reading of one bench.py JSON line (the last line of the file)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("N", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "G pairs/s %.2f" % (d["value"] / 1e9), "dominant:", r.get("kernel"), "frac %.3f" % r.get("frac", 0),
      "parity", d.get("parity_check"))
print("  build %.3f" % r.get("index_build_ms", 0), {k.split("(")[0].strip(): round(v["ms"], 3) for k, v in r.get("all_stages", {}).items()},
      "bin %.3f unbin %.3f" % (r.get("bin_ms", 0), r.get("unbin_ms", 0)))
if "e2e" in d and d["e2e"]:
    print("  e2e ms %.1f" % d["e2e"].get("ms_per_step", 0), "value %.3g" % d["e2e"]["value"], d["e2e"].get("split_ms"))
s = d.get("secondary")
if s:
    print("  config2: ms/step %.4f build %.4f count %.4f p1 %.4f p2 %.4f count frac %.3f" % (s["ms_per_step"], s["index_build_ms"], s["count_overlaps_ms"], s["pass1_ms"], s["pass2_ms"], s["count_overlaps_frac_of_hbm_peak"]))
c = d.get("config", {})
if c.get("exchange_ms_per_step") is not None:
    print("  exchange ms %.3f" % c["exchange_ms_per_step"], c.get("exchange_host_laps_ms"))
