"""First-contact diagnostic for the GPU box: run every conv-GEMM case on both implementations and
print the error table (does not stop at the first failure).  Not part of the product or the tests."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys
import time
import traceback
import zlib

import torch

sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
from gemm_ref import Case  # noqa: E402
from test_ops_gpu import CASES  # noqa: E402

print(torch.cuda.get_device_name(0), flush=True)
only = sys.argv[1:] or list(CASES)
for impl in (1, 0):
    for name in only:
        kw, tol = CASES[name]
        if impl == 1 and kw["m"] > 2000:
            continue
        try:
            t = time.time()
            case = Case(seed=zlib.crc32(name.encode()) % 1000, **kw)
            res = case.compare(case.run(impl=impl))
            print(f"impl={impl} {name:20s} {time.time() - t:5.1f}s", {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
        except Exception as e:
            print(f"impl={impl} {name:20s} EXC {e!r}", flush=True)
            traceback.print_exc()
            if "CUDA" in repr(e) or "cuda" in repr(e):
                print("stopping: CUDA context is likely poisoned", flush=True)
                sys.exit(1)
