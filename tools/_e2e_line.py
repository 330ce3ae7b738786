import json, sys
e = json.loads(sys.stdin.read())["e2e"]
print(round(e["ms_per_step"], 2), "ms e2e; ceiling", round(e["pcie_ceiling_ms"], 2),
      [(x["step"][0], x["arm"][0], x["kernel_ms"], x["copy_ms"], x["wall_ms"]) for x in e["per_solve_rank0"]])
