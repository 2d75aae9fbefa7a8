import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ldpc_3gpp_matlab_b200 import capi
from oracle import oracle as O
from conftest import make_llr

dt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bg, Z, B = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
et = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
rng = np.random.default_rng(1)
d = O.dims(bg, Z)
info, llr = make_llr(O, bg, Z, B, d["N"] // 2 * 2, 1.0, rng)
if dt == 2:   # the sum-product (reference algorithm) kernel against oracle B
    h = capi.Handle(bg, Z, 6, et, algorithm=capi.ALG_BP)
    out = h.decode(llr.astype(np.float64), want_soft=True)
    ref = O.decode_bp(bg, Z, llr.astype(np.float64), 6, early_term=et, want_app=True)
    print("hard equal", (out["hard"] == ref["hard"]).all(), "app close", np.allclose(out["app"], ref["app"], rtol=1e-9, atol=1e-9),
          "iters", (out["iters"] == ref["iters"]).all(), "ok", (out["parity_ok"] == ref["parity_ok"]).all())
    sys.exit(0)
h = capi.Handle(bg, Z, 6, et, llr_dtype=dt)
out = h.decode(llr, want_soft=True)
ref = O.decode_nms(bg, Z, llr, 6, early_term=et, f16=bool(dt))
print("hard equal", (out["hard"] == ref["hard"]).all(), "app equal", (out["app"].view(np.uint32) == ref["app"].view(np.uint32)).all(),
      "iters", (out["iters"] == ref["iters"]).all(), "ok", (out["parity_ok"] == ref["parity_ok"]).all())
if not (out["app"].view(np.uint32) == ref["app"].view(np.uint32)).all():
    bad = np.argwhere(out["app"].view(np.uint32) != ref["app"].view(np.uint32))
    print("first mismatches", bad[:10], out["app"][tuple(bad[0])], ref["app"][tuple(bad[0])], "count", len(bad))
