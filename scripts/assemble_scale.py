"""gpurun_out/r2_scale_N{1,2,4,8}.jsonl (scripts/scale_sweep.sh) -> profiles/r2_scale_<workload>.json: value, step time, end-to-end value,
exchange statistics and weak-scaling efficiency (value_N / (N * value_1)) per GPU count."""
import json
import sys
from pathlib import Path
root = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
rows = {}
for n in (1, 2, 4, 8):
    p = root / "gpurun_out" / f"{tag}_scale_N{n}.jsonl"
    if not p.exists():
        continue
    for ln in p.read_text().splitlines():
        try:
            d = json.loads(ln)
        except Exception:
            continue
        key = "snunet" if "SNUNet" in d["metric"] else "changeformer" if "ChangeFormer" in d["metric"] else \
              "floodvit-upernet" if "UPerNet" in d["metric"] else "floodvit" if "FloodViT" in d["metric"] else "siam-conc"
        rows.setdefault(key, {})[n] = d
for key, by_n in rows.items():
    base = by_n.get(1, {}).get("value")
    out = {"workload": by_n[min(by_n)]["config"]["workload"], "metric": by_n[min(by_n)]["metric"], "unit": "patches/s", "scaling": "weak",
           "note": "each N is a separate gpurun box (N GPUs of one 8xB200 node); box-to-box GPU variance is ~3 %; "
                   "efficiency = value_N / (N * value_1); ms_exposed = step time minus the same step with the exchange disabled (ranks free-running), "
                   "i.e. transfer that is not hidden plus rank skew", "points": []}
    for n in sorted(by_n):
        d = by_n[n]
        out["points"].append({"n_gpus": n, "value": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
                              "efficiency": (d["value"] / (n * base)) if base else None, "allreduce": d.get("allreduce"), "clocks": d.get("clocks")})
    (root / "profiles" / f"{tag}_scale_{key}.json").write_text(json.dumps(out, indent=1))
    print(key, [(p["n_gpus"], round(p["value"]), round(p["efficiency"], 3) if p["efficiency"] else None) for p in out["points"]])
