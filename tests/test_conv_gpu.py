"""GPU parity of the tcgen05/TMA convolution kernel, one case per layer kind, through the C ABI.

Checker: CPU fp32 torch ops on the same fp16-rounded operands (the reference's own call sites), and the
SIMT global-memory checker kernel for bisecting. Tolerance: the kernel accumulates in fp32 and rounds the
output to fp16 once, so |err| <= 1 fp16 ulp of the output magnitude (+ fp16 rounding of pre-summed weights).
"""

import ctypes

import pytest
import torch

import conv_cases
from robosat_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", range(len(conv_cases.default_cases(None))))
def test_conv_case_matches_cpu_reference(i, cuda_device):
    lib = _lib.load()
    case = conv_cases.default_cases(cuda_device)[i]()
    ref = case.ref()
    stream = _lib.current_stream_ptr()
    plan = ctypes.c_void_p()
    _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan_create")
    try:
        for _ in range(3):  # re-running the same plan must give the same answer (pipeline state resets)
            case.out.zero_()
            _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
            torch.cuda.synchronize()
            got = case.result()
            tol = 6e-3 * max(1.0, ref.abs().max().item() / 8)
            assert (got - ref).abs().max().item() <= tol, case.name
        # bisecting aid: the SIMT checker agrees with the tensor-core path to within output rounding
        tc = case.result()
        case.out.zero_()
        _lib.check(lib.rsb_conv_run_simt_check(ctypes.byref(case.desc), stream), "simt")
        torch.cuda.synchronize()
        assert (case.result() - tc).abs().max().item() <= tol
    finally:
        lib.rsb_conv_plan_destroy(plan)


@pytest.mark.parametrize("i", range(len(conv_cases.split_cases(None))))
def test_split_conv_case_matches_fp64_reference(i, cuda_device):
    """Strict precision (hi/lo fp16 operand planes, 3 MMAs per K step, fp32 accumulate in tensor memory) against a float64
    reference of the same operation on the values the planes represent. Tolerance: 2e-5 of the output range. (The tensor
    core truncates its fp32 accumulator after every MMA, so a single accumulation chain over K = 9216 was 4e-5 off; the plan
    cuts long K loops into chunks that the epilogue adds in round-to-nearest fp32 -- rsb_conv_desc.kchunk.)"""
    lib = _lib.load()
    case = conv_cases.split_cases(cuda_device)[i]()
    ref = case.ref()
    stream = _lib.current_stream_ptr()
    plan = ctypes.c_void_p()
    _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan_create")
    try:
        outs = []
        for _ in range(2):
            case.out.zero_()
            _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
            torch.cuda.synchronize()
            outs.append(case.out.clone())
        assert torch.equal(outs[0], outs[1]), "split conv is not deterministic"
        got = case.result()
        scale = max(1.0, ref.abs().max().item())
        err = (got - ref).abs().max().item()
        print("%s: max|err| %.3e (%.2e of range)" % (case.name, err, err / scale))
        assert err <= 2e-5 * scale, (case.name, err)
        # the SIMT checker (plain fp32 FMAs on hi + lo) agrees
        case.out.zero_()
        _lib.check(lib.rsb_conv_run_simt_check(ctypes.byref(case.desc), stream), "simt")
        torch.cuda.synchronize()
        assert (case.result() - got).abs().max().item() <= 4e-5 * scale
    finally:
        lib.rsb_conv_plan_destroy(plan)


def test_split_prepass_and_maxpool(cuda_device):
    """strict-precision helpers: the pre-pass writes half(x) and half(x - half(x)); the max pool of pairs == max of the sums"""
    import emulate
    from robosat_b200 import synth

    lib = _lib.load()
    u8 = synth.make_tiles_u8(2, 64, seed=5)
    ref = emulate.prepass_s2d_split_cpu(synth.normalize_tiles(u8))
    dst = torch.zeros(2, 2, 32, 36, 16, dtype=torch.float16, device=cuda_device)
    mean = (ctypes.c_float * 3)(*synth.IMAGENET_MEAN)
    std = (ctypes.c_float * 3)(*synth.IMAGENET_STD)
    u8d = u8.to(cuda_device)
    _lib.check(lib.rsb_prepass_s2d_split(u8d.data_ptr(), 1, dst.data_ptr(), dst.numel() // 2, 2, 64, 64, mean, std, _lib.current_stream_ptr()), "prepass")
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), ref)
    g = torch.Generator().manual_seed(3)
    for (N, H, W, C, k, s, p) in [(2, 32, 32, 64, 3, 2, 1), (3, 8, 8, 2048, 2, 2, 0)]:
        x = torch.randn((N, C, H, W), generator=g)
        hi = x.half()
        lo = (x - hi.float()).half()
        val = hi.float() + lo.float()
        ref = torch.nn.functional.max_pool2d(val, k, s, p)
        src = torch.stack([hi, lo]).permute(0, 1, 3, 4, 2).contiguous().to(cuda_device)
        dst = torch.zeros(2, N, ref.shape[2], ref.shape[3], C, dtype=torch.float16, device=cuda_device)
        _lib.check(lib.rsb_maxpool_nhwc_split(src.data_ptr(), src.numel() // 2, dst.data_ptr(), dst.numel() // 2, N, H, W, C, k, s, p,
                                              _lib.current_stream_ptr()), "maxpool_split")
        torch.cuda.synchronize()
        d = dst.cpu()
        assert torch.equal((d[0].float() + d[1].float()).permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_conv_multi_wave_persistent_schedule(block_n, cuda_device):
    """More tiles than SMs: every CTA loops over several tiles and both TMEM accumulator stages are reused."""
    lib = _lib.load()
    case = conv_cases.conv_case("3x3", 8, 64, 64, 64, 256, cuda_device, seed=21, block_n=block_n)
    plan = ctypes.c_void_p()
    _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan_create")
    _lib.check(lib.rsb_conv_run(plan, _lib.current_stream_ptr()), "conv_run")
    torch.cuda.synchronize()
    ref = case.ref()
    assert (case.result() - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item() / 8)
    lib.rsb_conv_plan_destroy(plan)


_PAIR_CASES = [
    lambda d: conv_cases.conv_case("1x1", 1, 16, 32, 256, 128, d, seed=2, residual=True, block_n=128),
    lambda d: conv_cases.conv_case("3x3", 1, 24, 40, 128, 256, d, seed=4, block_n=256),
    lambda d: conv_cases.conv_case("3x3", 8, 64, 64, 64, 256, d, seed=21, block_n=128),   # multi-wave
    lambda d: conv_cases.conv_case("3x3", 5, 8, 8, 128, 256, d, seed=22, block_n=256),    # odd number of spatial tiles
    lambda d: conv_cases.conv_case("1x1", 3, 8, 8, 512, 2048, d, seed=7, residual=True, block_n=256),
    lambda d: conv_cases.conv_case("3x3s2", 2, 32, 32, 128, 128, d, seed=5, block_n=128),
    lambda d: conv_cases.decoder_case(3, 4, 4, [256, 256], 256, d, seed=10, block_n=128),
]


@pytest.mark.parametrize("i", range(len(_PAIR_CASES)))
def test_conv_cta_pair_matches_single_cta(i, cuda_device):
    """cta_group::2 schedule (two SMs per 256-row tile pair, half a weight tile each) == one CTA per tile, bit for bit."""
    lib = _lib.load()
    case = _PAIR_CASES[i](cuda_device)
    stream = _lib.current_stream_ptr()
    outs = []
    for pair in (0, 1):
        case.desc.cta_pair = pair
        plan = ctypes.c_void_p()
        _lib.check(lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)), "plan_create")
        try:
            for _ in range(2):
                case.out.zero_()
                _lib.check(lib.rsb_conv_run(plan, stream), "conv_run")
                torch.cuda.synchronize()
            outs.append(case.out.clone())
        finally:
            lib.rsb_conv_plan_destroy(plan)
    ref = case.ref()
    assert (case.result() - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item() / 8), case.name
    assert torch.equal(outs[0], outs[1]), case.name


def test_prepass_u8_matches_reference_transform(cuda_device):
    """uint8 NHWC -> normalised fp16 s2d == ToTensor + Normalize (predict.py:71-73) then fp16 rounding."""
    import emulate
    from robosat_b200 import synth

    lib = _lib.load()
    u8 = synth.make_tiles_u8(2, 64, seed=5)
    ref = emulate.prepass_s2d_cpu(synth.normalize_tiles(u8))
    dst = torch.zeros(2, 32, 36, 16, dtype=torch.float16, device=cuda_device)
    mean = (ctypes.c_float * 3)(*synth.IMAGENET_MEAN)
    std = (ctypes.c_float * 3)(*synth.IMAGENET_STD)
    u8d = u8.to(cuda_device)
    _lib.check(lib.rsb_prepass_s2d(u8d.data_ptr(), 1, dst.data_ptr(), 2, 64, 64, mean, std, _lib.current_stream_ptr()), "prepass")
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), ref)


def test_maxpool_matches_torch(cuda_device):
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    for (N, H, W, C, k, s, p) in [(2, 32, 32, 64, 3, 2, 1), (3, 8, 8, 2048, 2, 2, 0)]:
        x = torch.randn((N, C, H, W), generator=g).half()
        ref = torch.nn.functional.max_pool2d(x.float(), k, s, p).half()
        src = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
        dst = torch.zeros(N, ref.shape[2], ref.shape[3], C, dtype=torch.float16, device=cuda_device)
        _lib.check(lib.rsb_maxpool_nhwc(src.data_ptr(), dst.data_ptr(), N, H, W, C, k, s, p, _lib.current_stream_ptr()), "maxpool")
        torch.cuda.synchronize()
        assert torch.equal(dst.cpu().permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("i", range(len(conv_cases.row_cases(None))))
def test_row_conv_case_matches_cpu_reference(i, cuda_device):
    """line-buffer kernel (shared-memory row ring, shifted operand windows) against the same CPU references"""
    from robosat_b200.engine import RowConvOp

    case = conv_cases.row_cases(cuda_device)[i]()
    ref = case.ref()
    op = RowConvOp(case.name, case.desc)
    tol = 6e-3 * max(1.0, ref.abs().max().item() / 8)
    for _ in range(2):
        case.out.zero_()
        op.run(_lib.current_stream_ptr())
        torch.cuda.synchronize()
        assert (case.result() - ref).abs().max().item() <= tol, case.name


@pytest.mark.parametrize("i", range(len(conv_cases.row_split_cases(None))))
def test_row_split_case_matches_fp64_reference(i, cuda_device):
    """strict-precision line-buffer kernel (dec5 + final): hi/lo row ring, [W_hi | W_lo] in one MMA + A_lo W_hi in a second"""
    from robosat_b200.engine import RowConvOp

    case = conv_cases.row_split_cases(cuda_device)[i]()
    ref = case.ref()
    op = RowConvOp(case.name, case.desc)
    scale = max(1.0, ref.abs().max().item())
    for _ in range(2):
        case.out.zero_()
        op.run(_lib.current_stream_ptr())
        torch.cuda.synchronize()
        err = (case.result() - ref).abs().max().item()
        print("%s: max|err| %.3e (%.2e of range)" % (case.name, err, err / scale))
        assert err <= 1e-5 * scale, case.name
