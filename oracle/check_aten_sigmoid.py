#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Exhaustive pin of oracle/csrc/aten_sigmoid.c: torch.sigmoid (CPU) vs the C restatement for
all 2^32 fp32 bit patterns.

    python -m oracle.check_aten_sigmoid           # ~70 s on 8 cores; prints the number of mismatching inputs (0 expected)
"""
import sys
import time

import numpy as np
import torch

from .build_native import aten_sigmoid


def main():
    t0 = time.time()
    bad = 0
    chunk = 1 << 26
    for start in range(0, 1 << 32, chunk):
        x = np.arange(start, start + chunk, dtype=np.uint64).astype(np.uint32).view(np.float32)
        ref = torch.sigmoid(torch.from_numpy(x)).numpy()
        got = aten_sigmoid(x)
        neq = (ref.view(np.uint32) != got.view(np.uint32)) & ~(np.isnan(ref) & np.isnan(got))
        if neq.any():
            i = int(np.nonzero(neq)[0][0])
            print(f"chunk {start:#x}: {int(neq.sum())} mismatches, first x={x[i]!r} torch={ref[i]!r} c={got[i]!r}")
            bad += int(neq.sum())
    print(f"torch {torch.__version__} ({torch.backends.cpu.get_cpu_capability()}): {bad} mismatching inputs of 2^32, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
