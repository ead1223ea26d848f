"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/mreserve_b200.h declares, descriptor structs agree in size with the C side, and argument
validation returns error codes (no compute kernels are launched here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mreserve_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    cdll = lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in the header but not exported"
    assert cdll.mr_abi_version() == lib.ABI_VERSION == 16
    assert lib.launch_count() >= 0


def test_binding_covers_header(lib):
    bound = set(lib.ENTRY_POINTS) | set(lib.WORKSPACE_FNS) | set(lib.MISC_FNS)
    assert set(_declared()) == bound


def test_struct_sizes_match_c(lib, tmp_path):
    """Compile a tiny C program against the public header and compare sizeof() with ctypes."""
    structs = {"MrGemmDesc": lib.GemmDesc, "MrAttentionDesc": lib.AttentionDesc, "MrLayerNormDesc": lib.LayerNormDesc,
               "MrCastDesc": lib.CastDesc, "MrFillClsDesc": lib.FillClsDesc, "MrWindowDesc": lib.WindowDesc,
               "MrAssembleDesc": lib.AssembleDesc, "MrRotaryDesc": lib.RotaryDesc, "MrNormalizeDesc": lib.NormalizeDesc,
               "MrSegmentPoolDesc": lib.SegmentPoolDesc, "MrContrastiveDesc": lib.ContrastiveDesc,
               "MrRowDotDesc": lib.RowDotDesc, "MrEncoderDesc": lib.EncoderDesc, "MrElementwiseDesc": lib.ElementwiseDesc,
               "MrColsumDesc": lib.ColsumDesc, "MrRotaryBwdDesc": lib.RotaryBwdDesc, "MrAdamWDesc": lib.AdamWDesc,
               "MrAdamWMultiDesc": lib.AdamWMultiDesc, "MrAdamWChunk": lib.AdamWChunk, "MrPatchifyDesc": lib.PatchifyDesc,
               "MrEncoderWeights": lib.EncoderWeights, "MrPoolWeights": lib.PoolWeights, "MrVisionDesc": lib.VisionDesc,
               "MrAudioDesc": lib.AudioDesc, "MrSpanDesc": lib.SpanDesc, "MrJointDesc": lib.JointDesc,
               "MrEmbedVideoDesc": lib.EmbedVideoDesc, "MrLogmelDesc": lib.LogmelDesc,
               "MrResizePatchifyDesc": lib.ResizePatchifyDesc}
    src = tmp_path / "sz.c"
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src.write_text(f'#include <stdio.h>\n#include "mreserve_b200.h"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert C.sizeof(structs[name]) == int(size), name


def test_argument_validation_without_gpu(lib):
    """Bad descriptors are rejected before anything touches the device."""
    cdll = lib.load()
    d = lib.GemmDesc(M=4, N=8, K=16, lda=12, ldd=8)  # row pitch smaller than K
    arr = (C.c_void_p * 5)(16, 16, None, None, None)
    out = (C.c_void_p * 1)(16)
    assert cdll.mr_gemm_epilogue(None, C.byref(d), arr, out, None, 0) == -1
    d = lib.GemmDesc(M=4, N=8, K=16, lda=16, ldd=8)
    arr = (C.c_void_p * 5)(None, 16, None, None, None)
    assert cdll.mr_gemm_epilogue(None, C.byref(d), arr, out, None, 0) == -4
    arr = (C.c_void_p * 5)(24, 16, None, None, None)  # misaligned A
    assert cdll.mr_gemm_epilogue(None, C.byref(d), arr, out, None, 0) == -2
    a = lib.AttentionDesc(B=1, T=4, H=100)
    arr = (C.c_void_p * 3)(16, None, None)
    assert cdll.mr_attention_fwd(None, C.byref(a), arr, out, None, 0) == -1
    assert cdll.mr_status_string(-2).decode().startswith("pointer")


def test_no_fallback_without_cuda(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from merlot_reserve_b200 import MerlotReserve
    from merlot_reserve_b200.synthetic import tiny_config
    model = MerlotReserve.from_config(tiny_config())
    with pytest.raises(lib.MrError):
        model.bind({})


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+[\w.]*oracle", re.M)
    for f in (ROOT / "merlot_reserve_b200").rglob("*.py"):
        assert not pat.search(f.read_text()), f


def test_wgrad_plan_fills_the_sms(lib):
    """Host logic only: every plan keeps tiles x splits within a whole number of waves that is >= 85 % full, or is
    a single partial wave of a small problem."""
    for o, i, rows in [(3072, 768, 30848), (768, 3072, 30848), (2304, 768, 30848), (768, 768, 30848), (3072, 768, 11904),
                       (768, 768, 6144), (768, 192, 11520), (1024, 4096, 73856), (4096, 1024, 73856), (3072, 1024, 73856)]:
        bn, s = lib.wgrad_plan(o, i, rows)
        assert bn in (128, 256) and (bn == 128 or i % 256 == 0) and 0 <= s <= 16 and s != 1
        items = -(-o // 128) * -(-i // bn) * max(s, 1)
        waves = -(-items // 148)
        assert items / (waves * 148) >= 0.85 or waves == 1, (o, i, rows, bn, s, items)
