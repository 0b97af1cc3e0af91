"""CPU-side checks: the C-ABI library loads and exports every symbol include/stfb200.h declares, argument
validation works without a device, the module tree reproduces the reference state_dict contract, and the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import stf_unet_b200 as S
from oracle import weights as W
from stf_unet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "stfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/stfb200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.EXPORTS) == syms
    assert lib.stfb_version() >= 100


def test_argument_validation_needs_no_device():
    lib = _lib.load()
    p = _lib.ConvParams()
    assert lib.stfb_conv2d(ctypes.byref(p), None) == -1
    assert b"null" in lib.stfb_last_error()
    assert lib.stfb_maxpool_fwd(None, None, 1, 4, 4, 4, 2, 2, 2, 2, 0, 0, None) == -1
    assert lib.stfb_ce_dice_fwd(None, None, None, None, 1, 2, 4, 1e-6, None) == -1
    # split-precision (bf16 x 3) entry points: channel counts must be multiples of 8; the operands belong to the tcgen05 family
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    assert lib.stfb_split_bf16x3(addr, addr, 4, 12, None) == -1 and b"multiple of 8" in lib.stfb_last_error()
    assert lib.stfb_split_bf16x3(None, None, 0, 64, None) == 0                      # empty input: nothing to do, no device needed
    assert lib.stfb_pack_weight_split(None, None, 8, 8, 3, 3, 1, 0, None) == -1
    p = _lib.ConvParams(x=addr, w=addr, y=addr, N=1, H=8, W=8, C1=64, C2=0, Ho=8, Wo=8, Cout=64, kh=3, kw=3, stride=1, pad=1, ldw=64,
                        mode=0, x_dtype=_lib.BF16X3, y_dtype=_lib.F32, impl=_lib.IMPL_SIMT)
    assert lib.stfb_conv2d(ctypes.byref(p), None) == -1 and b"tcgen05" in lib.stfb_last_error()
    p.impl, p.ldw = _lib.IMPL_TCGEN05, 9 * 64                                       # six K segments per channel: ldw too small
    assert lib.stfb_conv2d(ctypes.byref(p), None) == -1 and b"x6" in lib.stfb_last_error()
    p.ldw = 9 * 6 * 64
    assert lib.stfb_conv2d_tcgen05_supported(ctypes.byref(p)) == 1
    p.y_dtype = _lib.BF16                                                            # bf16x3 operands always produce fp32
    assert lib.stfb_conv2d_tcgen05_supported(ctypes.byref(p)) == 0
    assert lib.stfb_conv2d_wgrad(addr, addr, addr, 1, 8, 8, 64, 8, 8, 64, 0, 64, 3, 3, 1, 1, _lib.BF16X3, _lib.IMPL_SIMT, None, 0, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_device_is_an_error_not_a_fallback():
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    st = lib.stfb_cast(addr, 0, addr, 0, 16, None)
    assert st == -3 and b"no CPU fallback" in lib.stfb_last_error()
    m = S.UNet(1, 2, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.criterion({"out": torch.zeros(1, 2, 4, 4)}, torch.zeros(1, 4, 4, dtype=torch.long))


@pytest.mark.parametrize("kw", [dict(), dict(use_pk_maps=True), dict(in_channels=2, num_classes=3)])
def test_stf_state_dict_contract(kw):
    m = S.STFLSTMUNet(**kw)
    spec = W.stf_param_spec(kw.get("in_channels", 1), kw.get("num_classes", 2), kw.get("use_pk_maps", False))
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(n, tuple(s)) for n, s, _ in spec]
    assert all(v.dtype == (torch.int64 if k.endswith("num_batches_tracked") else torch.float32) for k, v in m.state_dict().items())
    assert not hasattr(m, "input_format")          # resolves to "time_sequence" (train_and_eval.py:10)
    sd = W.make_state_dict(spec)
    m.load_state_dict(sd)                           # strict


def test_unet_state_dict_contract():
    m = S.UNet(in_channels=8, num_classes=2, base_c=64)
    spec = W.unet_param_spec(8, 2, 64)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(n, tuple(s)) for n, s, _ in spec]
    assert S.UNet.input_format == "flat_channels"
    assert sum(p.numel() for p in m.parameters()) == 31046466


def test_param_counts_match_reference():
    assert sum(p.numel() for p in S.STFLSTMUNet().parameters()) == 27379746
    assert sum(p.numel() for p in S.UNet(1, 2, 64).parameters()) == 31042434


def test_criterion_argument_checks_and_no_cpu_fallback():
    """Every argument of the reference's criterion is accepted (loss_weight, dice, ignore_index); shape / class-count
    mismatches raise before any launch; CPU tensors raise (no fallback)."""
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        S.criterion({"out": torch.zeros(1, 2, 4, 4)}, torch.zeros(1, 4, 4, dtype=torch.long), ignore_index=255)
    with pytest.raises(ValueError, match="size mismatch"):
        S.ce_dice(torch.zeros(1, 2, 4, 4), torch.zeros(1, 8, 8, dtype=torch.long))
    with pytest.raises(ValueError, match="loss_weight"):
        S.criterion({"out": torch.zeros(1, 2, 4, 4)}, torch.zeros(1, 4, 4, dtype=torch.long), loss_weight=torch.ones(3))
    with pytest.raises(TypeError):
        S.ce_dice(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int32))
    # build_target mirrors dice_coefficient_loss.py:5-17 (ignored pixels carry ignore_index in every channel)
    from stf_unet_b200.loss import build_target
    t = torch.tensor([[[0, 1], [255, 1]]])
    bt = build_target(t, 2, 255)
    assert bt.shape == (1, 2, 2, 2) and bt[0, :, 1, 0].tolist() == [255.0, 255.0] and bt[0, :, 0, 1].tolist() == [0.0, 1.0]


def test_product_synthetic_generator_is_the_oracles_twin():
    """bench.py's own arm draws batches from stf_unet_b200.synthetic (never from oracle/); same recipe, same bits."""
    from stf_unet_b200.synthetic import synthetic_dce_batch, synthetic_dce_batch_u8
    for args in [(3, 4, 64, 64, 1234, True), (2, 2, 32, 48, 7, False)]:
        a, b = W.synthetic_dce_batch(*args), synthetic_dce_batch(*args)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    u8, tgt = synthetic_dce_batch_u8(2, 3, 32, 32, seed=5)
    x, tgt2 = synthetic_dce_batch(2, 3, 32, 32, seed=5)
    assert u8.dtype == torch.uint8 and u8.shape == (2, 3, 32, 32) and torch.equal(tgt, tgt2)
    # the 8-bit series is the float series quantised to 1/255 before normalisation
    back = (u8.float() / 255.0 - 0.709) / 0.127
    assert (back - x[:, :, 0]).abs().max() <= 0.5 / 255.0 / 0.127 + 1e-5


def test_volume_slice_ranges_partition_the_case():
    from stf_unet_b200.volume import slice_range
    for n, world in [(160, 8), (160, 1), (11, 3), (5, 8), (0, 2)]:
        ranges = [slice_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert slice_range(160, 3, 8) == (60, 80)
    with pytest.raises(ValueError):
        slice_range(10, 2, 2)


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The drop-in boundary is a C ABI: include/stfb200.h must compile as C99 and as C++, and a plain C program must link
    against libstfb200.so and reach the no-device error path (no torch, no Python in between)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "stfb200.h")
    subprocess.run([gcc, "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", hdr], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-x", "c++", "-fsyntax-only", hdr], check=True)
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "stfb200.h"\n'
                   'int main(void) {\n'
                   '  if (stfb_version() < 100) return 2;\n'
                   '  int st = stfb_maxpool_fwd(0, 0, 1, 4, 4, 4, 2, 2, 2, 2, 0, 0, 0);   /* null pointers: argument error */\n'
                   '  printf("%d %s\\n", st, stfb_last_error());\n'
                   '  return st == STFB_EINVAL ? 0 : 3;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "stf_unet_b200")
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-l:libstfb200.so", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "maxpool_fwd" in r.stdout


def test_augment_host_tables_and_draw_order_match_oracle():
    """The host half of the device-side augmentation (stf_unet_b200/augment.py): Pillow's resize coefficient tables, the nearest
    tables, the rotation matrices and the order of the random draws equal the oracle's (which is pinned to the reference's own
    transforms by tests/golden/make_golden_augment.py).  No GPU involved."""
    import random
    import numpy as np
    from oracle import augment_oracle as AO
    from stf_unet_b200 import augment as A
    for insz, out in [(256, 128), (256, 307), (256, 224), (256, 255), (200, 173), (312, 270), (150, 224)]:
        b, k = A._bilinear_tables(insz, out)
        bo, ko = AO.resize_coeffs(insz, out)
        assert np.array_equal(b, bo) and np.array_equal(k, ko), (insz, out)
        assert np.array_equal(A._nearest_table(out, insz), AO.nearest_table(out, insz))
    assert A._bilinear_tables(256, 256) == (None, None)
    for seed in range(20):
        aug = A.PairedAugment(train=True, rng=random.Random(seed))
        p = aug.draw((256, 256))
        assert p == {k: v for k, v in AO.draw_params(random.Random(seed)).items() if k != "crop"}
        if p["rot"]:
            assert A._rotate_matrix(p["angle"], p["rw"], p["rh"]) == AO.rotate_matrix(p["angle"], p["rw"], p["rh"])
    _, samples, tables = A.PairedAugment(train=True, rng=random.Random(5)).plan(3, 256, 256)
    assert len(bytes(samples)) == 3 * 120 and tables.dtype == np.int32 and samples[1].tab_off > 0
