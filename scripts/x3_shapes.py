"""Prepared-K/V (xattn_x3.cu) vs raw-K/V (xattn_tc5.cu / xattn_kernels.cu) call times for every SD-1.5 shape; L2 flushed.
Usage: python scripts/x3_shapes.py [B] [iters]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from x3_dev import run  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for (L, D) in ((4096, 40), (1024, 80), (256, 160), (64, 160)):
    run(B, L, iters, D=D)
for (L, D) in ((9216, 40), (2304, 80), (576, 160), (144, 160)):
    run(8, L, iters, D=D)
