#!/bin/bash
# round 2, tenth call: per-role stage stamps of the cooperative setup kernel
O=gpurun_out/r2_10
mkdir -p $O
GWAT_B200_LIB=$PWD/variants/coopprof/libgwat_b200.so python tools/setup_stage_profile.py > $O/setup_roles.json 2> $O/setup_roles.err
tail -3 $O/setup_roles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_10/setup_roles.json"))
for cfg, roles in d.items():
    print(cfg)
    for r, rows in roles.items():
        print("  %-9s" % r, "  ".join("%s:%d(@%d)" % (x["step"].split(" (")[0][:12], x["cycles"], x["at"]) for x in rows))
PY
