"""BASELINE config 4: dopri8 float64 on every DETEST problem (tests/DETEST/detest.py:8-315) replicated over a trailing
batch of 4096, rtol = atol in {1e-3, 1e-6, 1e-9, 1e-12}: NFE, time per solve on one B200 (captured step body inside the
device-side loop) and the RMS error against dopri5 @ 1e-12 exactly as run.py:37-47 computes it, next to the unmodified
reference's NFE / error for the single trajectory (tests/golden/detest.pt).  One JSON object per line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P            # noqa: E402
import torchdiffeq_b200 as tdq  # noqa: E402

DEV = torch.device("cuda:0")
DET = torch.load(os.path.join(ROOT, "tests", "golden", "detest.pt"), weights_only=False)


def main():
    tot = {}
    for name in P.DETEST_NAMES:
        f, y0, t0 = P.detest(name)
        yb = y0.unsqueeze(-1).repeat(*([1] * y0.dim()), 4096).to(DEV)
        t = torch.tensor([t0, 20.0], dtype=torch.float64, device=DEV)
        truth = DET[name + "/truth"]["y"]
        for tol in (1e-3, 1e-6, 1e-9, 1e-12):
            st = {}
            opts = {"graph": True, "cache": True}
            with torch.no_grad():
                for _ in range(2):
                    y = tdq.odeint(f, yb, t, method="dopri8", rtol=tol, atol=tol, options=opts, _stats=st)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    y = tdq.odeint(f, yb, t, method="dopri8", rtol=tol, atol=tol, options=opts, _stats=st)
                b.record()
                b.synchronize()
            ms = a.elapsed_time(b) / 3
            got = y[-1][..., 0].cpu()
            err = float(torch.sqrt(torch.mean((truth - got) ** 2)))
            ref = DET["%s/dopri8/%g" % (name, tol)]
            nfe = 2 + 13 * st["attempts"]
            print(json.dumps({"problem": name, "tol": tol, "nfe": nfe, "nfe_reference": ref["nfe"], "ms": round(ms, 3),
                              "rms_err_vs_dopri5_1e-12": err, "rms_err_reference": ref["err"],
                              "traj_per_s": round(4096 / ms * 1e3)}), flush=True)
            k = "%g" % tol
            tot.setdefault(k, [0, 0, 0.0])
            tot[k][0] += nfe
            tot[k][1] += ref["nfe"]
            tot[k][2] += ms
            tdq.clear_cache()
    print(json.dumps({"totals": {k: {"nfe": v[0], "nfe_reference": v[1], "ms": round(v[2], 2)} for k, v in tot.items()}}),
          flush=True)


if __name__ == "__main__":
    main()
