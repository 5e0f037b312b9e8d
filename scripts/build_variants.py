"""Build tuning variants of libevok.so (CPU box; nvcc cross-compiles).  python scripts/build_variants.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evotorch_b200 import build as B  # noqa: E402

VARIANTS = {
    "g_ldg": ("EVOK_GRAD_TMA_DEFAULT=0",),
    "gt_r4_s3_c4": ("EVOK_GRAD_TMA_ROWS=4", "EVOK_GRAD_TMA_STAGES=3", "EVOK_GRAD_TMA_CTAS_PER_SM=4"),
    "gt_r2_s8_c3": ("EVOK_GRAD_TMA_ROWS=2", "EVOK_GRAD_TMA_STAGES=8", "EVOK_GRAD_TMA_CTAS_PER_SM=3"),
    "gt_r8_s2_c3": ("EVOK_GRAD_TMA_ROWS=8", "EVOK_GRAD_TMA_STAGES=2", "EVOK_GRAD_TMA_CTAS_PER_SM=3"),
    "gt_r4_s2_c6": ("EVOK_GRAD_TMA_ROWS=4", "EVOK_GRAD_TMA_STAGES=2", "EVOK_GRAD_TMA_CTAS_PER_SM=6"),
    "gt_r2_s4_c6": ("EVOK_GRAD_TMA_ROWS=2", "EVOK_GRAD_TMA_STAGES=4", "EVOK_GRAD_TMA_CTAS_PER_SM=6"),
    "gt_r6_s3_c3": ("EVOK_GRAD_TMA_ROWS=6", "EVOK_GRAD_TMA_STAGES=3", "EVOK_GRAD_TMA_CTAS_PER_SM=3"),
    "so_unr1_minb6": ("EVOK_SAMPLEONLY_UNR=1", "EVOK_SAMPLEONLY_MINB=6"),
    "so_unr2_minb4": ("EVOK_SAMPLEONLY_UNR=2", "EVOK_SAMPLEONLY_MINB=4"),
    # round 2: occupancy of the fused sampler (71 registers -> 3 CTAs x 8 warps per SM; ncu: 28 % warps active, issue 63 %, XU 66 %)
    "se_minb4": ("EVOK_SAMPLE_MINB=4",),
    "se_unr1_minb4": ("EVOK_SAMPLE_MINB=4", "EVOK_SAMPLE_UNR=1"),
    "se_unr1_minb5": ("EVOK_SAMPLE_MINB=5", "EVOK_SAMPLE_UNR=1"),
    "se_t128_minb8": ("EVOK_SAMPLE_THREADS=128", "EVOK_SAMPLE_MINB=8"),
    "se_t128_minb7": ("EVOK_SAMPLE_THREADS=128", "EVOK_SAMPLE_MINB=7"),
    "se_t128_unr1_minb10": ("EVOK_SAMPLE_THREADS=128", "EVOK_SAMPLE_MINB=10", "EVOK_SAMPLE_UNR=1"),
    # measurement only: how much of the fused sampler is the counter-based RNG?  (never shipped: the product is Philox4x32-10)
    "philox7": ("EVOK_PHILOX_ROUNDS=7",),
    "philox4": ("EVOK_PHILOX_ROUNDS=4",),
}
only = sys.argv[1:]
for tag, defs in VARIANTS.items():
    if only and tag not in only:
        continue
    print(tag, B.build(defines=defs, tag=tag, verbose=False))
