#!/usr/bin/env python
"""Print the headline numbers of bench.py JSON lines read from stdin (development helper)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.get("kernel_time_share", {})
    print(sys.argv[1] if len(sys.argv) > 1 else "", f"{d['value'] / 1e9:.3f} G", f"{d['ms_per_step']:.3f} ms/step",
          "elem", round(k.get("element_ms", 0), 2), "surf", round(k.get("surface_flux_ms", 0), 2),
          "cfl", round(k.get("max_dt_ms", 0), 2), "halo", round(k.get("halo_pack_wait_mpiflux_ms", 0), 2),
          "frac", round(d.get("roofline", {}).get("frac", 0), 3), "sm_mhz", d.get("clocks", {}).get("sm_mhz"))
