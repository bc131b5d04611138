"""Reads a bench.py JSON line on stdin and prints its headline fields (helper for sweeps under gpurun)."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(" ".join(sys.argv[1:]), "value %.1f  e2e %.1f  ms/step %.2f  pool %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("host_threads_per_rank")))
