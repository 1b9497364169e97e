#!/usr/bin/env python
"""Host time line of bbcp_probe_lits_to_host on the C2 workload (BBCP_E2E_TRACE), for several chunk counts.
Usage: python tools/e2e_trace.py [chunks ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from intel_sat_solver_b200 import BatchedTopor, fileio  # noqa: E402
from intel_sat_solver_b200._lib import BatchInfo, Opts  # noqa: E402

nvars, ncls = 1_000_000, 4_200_000
inst = bench.make_instance(nvars, ncls, 1, "c2")
eng = BatchedTopor(device=0)
eng.AddClauses(fileio.read_bcnf(inst))
assert eng.Finalize() == 0
L, h = eng.lib, eng.handle
n = 2 * nvars
idx = np.arange(0, n, dtype=np.int64)
h_lits = torch.from_numpy(np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)).pin_memory()
o_flag = torch.empty(n + 8, dtype=torch.uint8).pin_memory()
o_off = torch.empty(n + 1, dtype=torch.int32).pin_memory()
o_lits = torch.empty(4 * n, dtype=torch.int32).pin_memory()
info = BatchInfo()
for flags, name in ((1 | 4 | 128, "compact"), (1 | 4, "full")):
    for mode in ("dma", "mirror", "pull"):
        os.environ["BBCP_E2E"] = mode
        for chunks in [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3]:
            if chunks:
                os.environ["BBCP_CHUNKS"] = str(chunks)
            else:
                os.environ.pop("BBCP_CHUNKS", None)
            opts = Opts(flags, 0, 0, 0, 0)
            secs = []
            for it in range(12):
                if it == 10:
                    os.environ["BBCP_E2E_TRACE"] = "1"
                L.bbcp_l2_flush(h, 0)
                L.bbcp_sync(h)
                t0 = time.perf_counter()
                rc = L.bbcp_probe_lits_to_host(h, h_lits.data_ptr(), n, C.byref(opts), o_flag.data_ptr(), o_off.data_ptr(), o_lits.data_ptr(), o_lits.numel(), C.byref(info))
                assert rc == 0
                torch.cuda.synchronize()
                secs.append(time.perf_counter() - t0)
            os.environ.pop("BBCP_E2E_TRACE", None)
            sys.stderr.flush()
            print(f"{name} {mode} chunks={chunks or 'default'}: median {1e6 * float(np.median(secs[3:])):.1f} us, min {1e6 * min(secs[3:]):.1f} us -> "
                  f"{n / float(np.median(secs[3:])) / 1e9:.2f} G probes/s", flush=True)
