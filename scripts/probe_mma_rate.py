"""tcgen05.mma issue-rate probe (csrc/probe.cu): cycles per MMA for the operand shapes the conv kernels use."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cet_pick_b200 import _lib

L = _lib.test_lib()
out = torch.zeros(148, device="cuda")
cases = [  # N, KC, sbo_a, a_step, ntap, ndst, note
    (32, 32, 512, 0, 1, 1, "N=32 dense"),
    (96, 32, 512, 0, 1, 1, "N=96 dense, same A"),
    (96, 32, 1536, 256, 9, 2, "N=96 march3d-like (sbo 1536, 9 taps, 2 tiles)"),
    (96, 32, 512, 64, 3, 2, "N=96 march2d-like KC=32"),
    (192, 64, 1024, 128, 3, 1, "N=192 march2d-like KC=64"),
    (128, 64, 1024, 0, 1, 2, "N=128 dense"),
    (128, 64, 2304, 128, 9, 2, "N=128 halo-like (sbo 2304)"),
    (256, 64, 1024, 0, 1, 2, "N=256 dense"),
    (64, 64, 1024, 0, 1, 2, "N=64 dense"),
]
for grid in (1, 148):
    for N, KC, sbo, step, ntap, ndst, note in cases:
        for _ in range(2):
            _lib.check(L.cetpick_probe_mma_rate(N, KC, sbo, step, ntap, ndst, 4096, out.data_ptr(), grid, _lib.stream_ptr()), "probe")
            torch.cuda.synchronize()
        c = out[:grid]
        print(f"grid={grid:3d} {note:52s} cycles/MMA min {c.min().item():7.1f} mean {c.mean().item():7.1f}  floor {128 * N / 256:5.1f}")

# CTA pair (tcgen05.mma.cta_group::2, M = 256): per-CTA operand traffic = own 128 rows of A + N/2 rows of B
out2 = torch.zeros(74, device="cuda")
for pairs in (1, 74):
    for N, KC, sbo, step, ntap, ndst, note in cases:
        if N % 32:
            continue
        for _ in range(2):
            _lib.check(L.cetpick_probe_mma_rate2(N, KC, sbo, step, ntap, ndst, 4096, out2.data_ptr(), pairs, _lib.stream_ptr()), "probe2")
            torch.cuda.synchronize()
        c = out2[:pairs]
        print(f"pairs={pairs:3d} cta_group::2 {note:44s} cycles/MMA min {c.min().item():7.1f} mean {c.mean().item():7.1f}  floor {128 * N / 256:5.1f}")
