"""Prints which fit baselines a bench line used (reference package or oracle port): python profiles/check_reference_arms.py LINE.json"""
import json,sys
d=json.loads(open(sys.argv[1]).read())
f=d["fit"]
print(f["lg_20x20"]["cpu_baseline"]["kind"], f["lg_20x20"].get("reference_cuda",{}).get("kind"), f["coevo_400x400"]["cpu_baseline"]["kind"], f["coevo_400x400"].get("reference_cuda",{}).get("seconds_end_to_end"), f.get("cpu_baseline_error"))
