"""Golden values for the host-side input readers (upsp-processing_b200/host/run_inputs.hpp), made with the
reference's own Python parsers (python/upsp/cam_cal_utils/parsers.py: read_tgts, read_wind_tunnel_data) from the
two sample files committed next to this script.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_inputs_golden.py   ->  tests/golden/inputs_golden.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/python")
from upsp.cam_cal_utils import parsers  # noqa: E402

tg = parsers.read_tgts(os.path.join(HERE, "sample.tgts"))
items = ("ALPHA", "BETA", "PHI", "PTOT", "TTF", "PS", "Q", "RNU", "TCAVG")
wtd = parsers.read_wind_tunnel_data(os.path.join(HERE, "sample.wtd"), items=items)
out = {
    "tgts": [{"idx": t["idx"], "xyz": [float(v) for v in t["tvec"][:, 0]], "size": t["size"], "name": t["name"]} for t in tg],
    "wtd": wtd,
}
with open(os.path.join(HERE, "inputs_golden.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
