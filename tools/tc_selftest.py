"""Runs msdf_tc_selftest for every variant (GPU box)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monosdf_b200 import _lib
torch.cuda.init()
torch.zeros(1, device="cuda")
ok = True
for v in range(_lib.lib().msdf_tc_selftest_count()):
    res = (ctypes.c_float * 2)()
    rc = _lib.lib().msdf_tc_selftest(v, res, None)
    if rc != 0:
        print("variant", v, "FAILED rc", rc, _lib.lib().msdf_last_error().decode()); ok = False
        break
    rel = res[0] / max(res[1], 1e-30)
    print("variant %d: max|err| %.4g  max|ref| %.4g  rel %.3g %s" % (v, res[0], res[1], rel, "OK" if rel < 6e-3 else "MISMATCH"))
    ok = ok and rel < 6e-3
sys.exit(0 if ok else 1)
