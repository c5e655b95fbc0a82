import sys, json
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print("value %.1f  e2e %.1f  single %.2f ms  sor_l0 %.0f GB/s (frac %.2f)  launches/solve %d" % (
            d["value"], d["e2e"]["value"], d.get("single_pair_latency_ms", 0), d["roofline"]["achieved"], d["roofline"]["frac"], d["launches_per_solve"]))
