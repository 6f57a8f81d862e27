"""Stand-alone diagnostic for the tcgen05 GEMM (not a pytest): prints error statistics and a coarse
error map so one GPU run localises descriptor / swizzle / epilogue mistakes."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
print(torch.cuda.get_device_name(0), torch.version.cuda)


def run(m, n, k, pattern="rand"):
    g = torch.Generator().manual_seed(0)
    if pattern == "rand":
        a = torch.randn(m, k, generator=g)
        w = torch.randn(n, k, generator=g) / math.sqrt(k)
    elif pattern == "eye":            # out[i, j] = a[i, j] for j < k : exposes row/col permutations
        a = torch.randn(m, k, generator=g)
        w = torch.eye(n, k)
    a, w = a.half().to(dev), w.half().to(dev)
    try:
        out = ops.gemm(a, w)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"gemm {m}x{n}x{k} [{pattern}] raised: {e}")
        return False
    ref = a.float() @ w.float().t()
    err = (out.float() - ref).abs()
    ok = bool((err <= 1e-4 + 1e-3 * ref.abs()).all())
    print(f"gemm {m}x{n}x{k} [{pattern}] tile_n={ops.gemm_tile_n(m, n, k)} max_err={float(err.max()):.4e} "
          f"mean_err={float(err.mean()):.4e} ref_rms={float(ref.pow(2).mean().sqrt()):.3f} ok={ok}")
    if not ok:
        # coarse map: fraction of bad elements per (32-row, 16-col) cell of the first tile
        bad = (err > 1e-4 + 1e-3 * ref.abs()).float()
        mm, nn = min(m, 128), min(n, 128)
        cells = bad[:mm, :nn]
        rows = []
        for r0 in range(0, mm, 32):
            rows.append(" ".join(f"{float(cells[r0:r0 + 32, c0:c0 + 16].mean()):.2f}" for c0 in range(0, nn, 16)))
        print("  bad-fraction map (32 rows x 16 cols cells):\n   " + "\n   ".join(rows))
        print("  out[0,:8] ", out[0, :8].float().tolist())
        print("  ref[0,:8] ", ref[0, :8].tolist())
        print("  out[1,:8] ", out[1, :8].float().tolist())
        print("  ref[1,:8] ", ref[1, :8].tolist())
    return ok


results = []
for shape in [(128, 64, 64), (128, 128, 64), (128, 160, 64), (128, 256, 64), (128, 64, 128), (128, 64, 512),
              (256, 128, 64), (8, 64, 64), (128, 64, 80), (8192, 320, 320), (8192, 2560, 320), (512, 1280, 1280)]:
    for pat in (("eye", "rand") if shape[2] <= 128 and shape[0] <= 256 else ("rand",)):
        results.append(run(*shape, pattern=pat))
print("ALL OK" if all(results) else f"FAILURES: {results.count(False)}/{len(results)}")
