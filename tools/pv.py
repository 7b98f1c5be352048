import json, sys
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        print("value %.1f pairs/s  %.1f ms/step  e2e %.1f  launches %s  power %s W  clocks %s  frac %.3f" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d["clocks"].get("power_w_max"),
            d["clocks"].get("sm_mhz"), d["roofline"]["frac"]))
