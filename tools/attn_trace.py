"""Phase breakdown of the tcgen05 attention kernel's softmax warps from clock64() stamps
(mr_debug_attention_trace): per key block, cycles spent waiting for S, loading S from TMEM, row max + mask,
scale + ex2 + pack, waiting for the previous PV, storing P.   python tools/attn_trace.py [vision|joint]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merlot_reserve_b200 import _lib as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "vision"
B, T, H, masked = {"vision": (1024, 241, 768, False), "joint": (64, 1120, 768, True),
                   "large_vision": (256, 577, 1024, False), "audio": (3072, 31, 768, False), "large_joint": (16, 2464, 1024, True)}[which]
torch.cuda.set_device(0)
lib = L.load()
lib.mr_debug_attention_trace.argtypes = [C.c_void_p]
qkv = torch.randn(B * T, 3 * H, device="cuda", dtype=torch.bfloat16)
qkv[:, :H] *= 0.125                                # the QKV GEMM epilogue hands attention q / sqrt(64)
valid = None
if masked:
    valid = torch.ones(B, T, dtype=torch.uint8, device="cuda")
    valid[:, 13:160] = 0
    valid = valid.reshape(-1)
for _ in range(2):
    L.attention(qkv, B, T, H, valid=valid)
buf = torch.zeros(2 * 64 * 8, dtype=torch.int64, device="cuda")
lib.mr_debug_attention_trace(C.c_void_p(buf.data_ptr()))
L.attention(qkv, B, T, H, valid=valid)
torch.cuda.synchronize()
lib.mr_debug_attention_trace(None)
t = buf.cpu().view(2, 64, 8)
names = ["wait S", "ld S", "max/mask", "ffma+ex2+pack", "wait PV(j-1)", "rescale O", "st P + arrive"]
nkb = (T + 127) // 128
if T <= 64:
    nkb = 1
for wg in range(2):
    print(f"tile {wg}: stamps relative to the first; one row per key block (block index within the CTA's stream)")
    t0 = int(t[wg, 0, 0])
    for n in range(min(64, 4 * nkb)):
        r = [int(x) for x in t[wg, n]]
        if r[0] == 0:
            break
        d = [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], (r[5] - r[4]) if r[5] >= r[4] else 0,
             (r[6] - max(r[5], r[4])), r[7] - r[6]]
        nxt = int(t[wg, n + 1, 0]) if n + 1 < 64 and int(t[wg, n + 1, 0]) else r[7]
        print(f"  blk {n:2d} (j={n % nkb}) start {r[0] - t0:7d} total {r[7] - r[0]:5d} gap-to-next {nxt - r[7]:5d} | " +
              " ".join(f"{nm}={v}" for nm, v in zip(names, d)))
