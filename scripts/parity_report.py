"""Prints the measured parity of the CUDA path against the reference's golden vectors (tests/golden/):
max-abs and RMS error per case.  Run on a GPU box:  python scripts/parity_report.py [--json out.json]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from tests.helpers import err, load_case  # noqa: E402
from tests.test_gpu_parity import _model  # noqa: E402

CASES = [("kat_randinit_newt", "randinit", False), ("kat_randinit_fast", "randinit", True),
         ("kat_vn_newt", "vn", False), ("kat_vn_fast", "vn", True), ("kat_fl_newt", "fl", False),
         ("kat_fl_fast", "fl", True), ("kat_tpt_newt", "tpt", False), ("kat_tpt_fast", "tpt", True),
         ("small_vn_newt", "vn", False), ("small_vn_fast", "vn", True), ("min_randinit_newt", "randinit", False)]


def main():
    rows = []
    for case, tag, fast in CASES:
        m, _ = _model(tag, fast)
        c = load_case(case)
        with torch.no_grad():
            y = m(c["f0"].cuda(), c["control"].cuda(), phase_shift=c["u_phase"].cuda(), noise=c["noise"].cuda())
        e = err(y, c["out"])
        rows.append({"case": case, "max_abs": e[0], "rms": e[1], "signal_max": float(c["out"].abs().max()),
                     "signal_rms": float(c["out"].pow(2).mean().sqrt())})
        print("%-22s max|err| %.3e  rms %.3e   (signal max %.3f rms %.3f)" % (case, e[0], e[1], rows[-1]["signal_max"],
                                                                             rows[-1]["signal_rms"]))
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
