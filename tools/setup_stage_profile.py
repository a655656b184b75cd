"""Kernel experiment: where the cooperative setup kernel spends its cycles, per role (lane 0 of every warp of block 0, clock64).
Needs a library built with -DGWAT_SETUP_PROFILE (tools/build_variant.sh coopprof -DGWAT_SETUP_PROFILE) selected by GWAT_B200_LIB."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from gw_analysis_tools_b200 import engine, workloads  # noqa: E402

STEPS = ["start", "repack / load", "step 1 (own stages)", "barrier A", "step 2 (carrier halves)", "barrier B", "step 3 (time-shift samples, ga/gb)",
         "barrier C", "step 4 (spline) + barrier D", "validity + barrier E", "write-out"]


def main():
    lib = engine.load_library()
    fn = lib.gwat_b200_debug_setup_stamps
    out = {}
    for cfg in (1, 2, 4):
        wl = workloads.make(cfg, L=2048)
        ctx = engine.Context(0)
        ctx.set_network(wl.detectors, wl.f, wl.psd, np.zeros((wl.D, wl.L), dtype=complex))
        for _ in range(3):
            ctx.loglike_mcmc_batch(wl.method, wl.params, wl.gmst, wl.T_segment, wl.mod)
        st = (C.c_longlong * 64)()
        assert fn(st) == 0
        st = np.array(list(st)).reshape(4, 16)
        t0 = st[:, 0][st[:, 0] > 0].min()
        roles = {}
        for r, name in enumerate(["phase", "amp", "detector", "twist"]):
            if st[r, 0] == 0:
                continue
            row, prev = [], st[r, 0]
            for i in range(1, 11):
                if st[r, i] > 0:
                    row.append({"step": STEPS[i], "cycles": int(st[r, i] - prev), "at": int(st[r, i] - t0)})
                    prev = st[r, i]
            roles[name] = row
        out[wl.name] = roles
        ctx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
