import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  ms/step %.3f  roofline %.1f%%  e2e %s  clocks %s" % (d["value"], d["unit"], d["ms_per_step"], d.get("pct_hbm_roofline", 0),
      d["e2e"] and round(d["e2e"]["value"], 1), d["clocks"]))
for k in d["roofline"]["kernels"]:
    print("   %-50s %8.3f ms x %d  share %.3f" % (k["name"][:50], k["ms_per_launch"], k["launches"], k["share"]))
print("   comm %.3f ms/step  kernels %.3f ms/step  checksum %s  mass %s" % (d.get("comm_ms_per_step", -1), d.get("kernel_ms_per_step", -1), d.get("checksum"), d.get("mass")))
