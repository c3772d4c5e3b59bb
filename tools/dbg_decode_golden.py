"""Debug helper: decode every golden 'small' stream through the QB3.h API, printing the case before each call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PRODUCT_SO, QB3Lib, golden_cases
P = QB3Lib(PRODUCT_SO, 256)
for compat in ("1", None):
    if compat: os.environ["QB3_REF_COMPAT"] = compat
    else: os.environ.pop("QB3_REF_COMPAT", None)
    for case in golden_cases():
        if case["kind"] != "small":
            continue
        print(compat, case["name"], flush=True)
        P.decode(bytes.fromhex(case["stream"]))
print("done")
