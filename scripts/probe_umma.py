"""Which shifted/strided UMMA descriptors read a TMA-swizzled tile correctly?  (GPU only)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cet_pick_b200 import _lib
L = _lib.test_lib()
R = 384
for KC in (64, 32, 16):
    g = torch.Generator(device="cuda").manual_seed(KC)
    A = torch.randn(R, KC, device="cuda", generator=g).bfloat16()
    B = torch.randn(32, KC, device="cuda", generator=g).bfloat16()
    full = A.float() @ B.float().t()           # (R, 32)
    rowb = KC * 2
    for r0 in (0, 8, 16, 1, 2, 3, 4, 5, 10):
        for pitch_rows in (8, 10, 12, 16, 24):
            ok = []
            for bo in range(8):
                out = torch.full((128, 32), float("nan"), device="cuda")
                _lib.check(L.cetpick_probe_umma(A.data_ptr(), R, B.data_ptr(), KC, r0, pitch_rows * rowb, bo,
                                                out.data_ptr(), None), "probe")
                torch.cuda.synchronize()
                rows = torch.tensor([r0 + (m // 8) * pitch_rows + (m % 8) for m in range(128)], device="cuda")
                if rows.max() >= R:
                    continue
                if (out - full[rows]).abs().max().item() < 1e-2:
                    ok.append(bo)
            print(f"KC={KC} r0={r0:2d} group pitch={pitch_rows:2d} rows: base_offset values that match: {ok}")
