"""One-line digest of a bench.py JSON line.  python scripts/show_bench.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
x = d.get("extra", {}).get("c4_pinhole", {})
print("N=%d  C2 device %.3f G rays/s (%.4f ms)  e2e %.4f ms (pageable %.4f)  find alone %.4f ms  stage %s | C4 %.4f ms, find %.4f ms, e2e %.4f ms | n_meas %s"
      % (d["n_gpus"], d["value"] / 1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_scan_ms_per_step", 0.0), d["stage_ms"]["find_alone"],
         {k: round(v, 4) for k, v in d["stage_ms"].items()}, x.get("ms_per_step", 0.0), x.get("find_alone_ms", 0.0), x.get("e2e_ms_per_step", 0.0), d["result_check"]["n_meas"]))
