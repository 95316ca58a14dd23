"""print the headline fields of a bench.py JSON line read from stdin"""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d["e2e"]
print(d["value"], d["ms_per_step"], d.get("host_step_interval_ms"), "| e2e", e["value"], e["ms_per_step"], e.get("host_step_interval_ms"),
      "| warmup", d["warmup"], "+", d["config"].get("settle_warmup_steps"), "clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("sm_mhz_min"), d["clocks"].get("reasons"),
      "| variants", d.get("variants"))
