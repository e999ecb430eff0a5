"""Probe tcgen05 descriptor behaviour on the device (see csrc/rsb_debug.cu); prints one JSON line per variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from robosat_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = torch.randn((64, 64), generator=g).half()
out = torch.zeros((128, 64), dtype=torch.float32, device=dev)
st = _lib.current_stream_ptr()


def run(**kw):
    out.zero_()
    rc = _lib.load_debug().rsb_debug_umma(kw["a"].data_ptr(), kw["a"].shape[0], kw["a"].shape[1], Bd.data_ptr(), out.data_ptr(), kw["mode"], kw["a_rows"],
                            kw["a_blocks"], kw.get("row_offset", 0), kw.get("base_offset", 0), kw.get("lbo", 0), kw.get("sbo", 0),
                            kw.get("k_step", 0), st)
    if rc:
        return {"error": _lib.last_error()}
    torch.cuda.synchronize()
    return out.cpu()


Bd = B.to(dev)
# (1) K-major shifted window
A = torch.randn((160, 64), generator=g).half()
Ad = A.to(dev)
for ro in (0, 1, 2, 3, 5, 8, 9, 17):
    ref = A[ro:ro + 128].float() @ B.float().t()
    for bo in sorted({0, ro & 7}):
        got = run(a=Ad, mode=0, a_rows=152, a_blocks=1, row_offset=ro, base_offset=bo)
        err = float((got - ref).abs().max()) if not isinstance(got, dict) else got
        print(json.dumps({"probe": "kmajor_shift", "row_offset": ro, "base_offset": bo, "max_err": err}), flush=True)
# (2) MN-major A: stored [64 k][128 m]
Amn = torch.randn((64, 128), generator=g).half()
Amnd = Amn.to(dev)
ref = Amn.float().t() @ B.float().t()
for lbo, sbo in ((8192, 1024), (1024, 8192)):
    got = run(a=Amnd, mode=1, a_rows=64, a_blocks=2, lbo=lbo, sbo=sbo, k_step=2048)
    err = float((got - ref).abs().max()) if not isinstance(got, dict) else got
    print(json.dumps({"probe": "mn_major", "lbo": lbo, "sbo": sbo, "max_err": err}), flush=True)
