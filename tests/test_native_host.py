"""CPU-side tests of libv2w: ABI surface, layer table, FLOP accounting, weight folding/packing
and the plan (executed by the numpy interpreter in plan_interp.py against the goldens).
No compute call reaches a GPU here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import fixtures, vec2wav_numpy
from wavthruvec_pytorch_b200 import Generator, native

import plan_interp
from helpers import dropin_in_state, hparams, load_golden, t

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_every_declared_symbol():
    header = (ROOT / "include" / "v2w.h").read_text()
    declared = set(re.findall(r"\b(v2w_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = native.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libv2w.so does not export {name}"
    assert declared == set(native.EXPORTED_SYMBOLS), declared ^ set(native.EXPORTED_SYMBOLS)
    assert lib.v2w_abi_version() == 1


@pytest.mark.parametrize("rb,feat,layers,flops", [
    (1, 768, 37, 118_888_448), (1, 1024, 37, 120_723_456),
    ('1', 768, 97, 325_326_848), ('1', 1024, 97, 327_161_856)])
def test_layer_table_and_flops(rb, feat, layers, flops):
    """Algorithmic FLOPs per input frame == SURVEY.md section 8(d) / BASELINE.md section 3."""
    gen = Generator(hparams(resblock=rb, num_wv_feat=feat))
    eng = gen.engine()
    ls = eng.layers()
    assert len(ls) == layers
    assert eng.flops_per_frame() == flops
    assert eng._lib.v2w_model_num_convs(eng._model) == len(gen.conv_modules())
    # polyphase tap sets of the five transposed convs: (k,u) -> taps per phase
    ups = [l for l in ls if l.role == native.ROLE_UPSAMPLE]
    assert [(l.kernel_size, l.up, l.num_taps) for l in ups] == [
        (11, 5, 3), (8, 4, 3), (8, 4, 3), (4, 2, 3), (4, 2, 3)]
    l0 = ups[0]
    per_phase = [sum((l0.tap_mask[j] >> p) & 1 for j in range(l0.num_taps)) for p in range(5)]
    assert sorted(per_phase) == [2, 2, 2, 2, 3]
    assert list(l0.tap_off[:3]) == [-1, 0, 1]


@pytest.mark.parametrize("bad", [
    dict(upsample_initial_channel=256), dict(num_wv_feat=100),
    dict(upsample_rates=[5, 4, 4, 2, 2], upsample_kernel_sizes=[10, 8, 8, 4, 4]),
    dict(resblock_kernel_sizes=[3, 4, 11]),
    dict(resblock='1', resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])])
def test_unsupported_hyperparameters_fail_loudly(bad):
    h = hparams(**bad)
    cfg = native.make_config(h)
    handle = C.c_void_p()
    rc = native.load().v2w_model_create(C.byref(cfg), C.byref(handle))
    assert rc == -1 and not handle.value
    assert native.last_error()


def test_pack_matches_numpy_fold():
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g)
    eng = gen.engine()
    packed = eng.pack_host().numpy()
    sd = vec2wav_numpy.state_dict_to_numpy(gen)
    names = ["conv_pre"] + [f"ups.{i}" for i in range(5)]
    for rb in range(15):
        names += [f"resblocks.{rb}.convs.0", f"resblocks.{rb}.convs.1"]
    names.append("conv_post")
    for L in eng.layers():
        name = names[L.conv_index]
        w = vec2wav_numpy.effective_weight(sd, name, np.float64)
        w32 = packed[L.w32_off: L.w32_off + 4 * L.num_taps * L.cin * L.n_total].view(
            np.float32).reshape(L.num_taps, L.cin, L.n_total)
        bias = packed[L.bias_off: L.bias_off + 4 * L.cout].view(np.float32)
        np.testing.assert_array_equal(bias, sd[name + ".bias"])
        if L.role != native.ROLE_UPSAMPLE:
            # Conv1d [Cout,Cin,K]: tap j <-> kernel index j
            want = w.transpose(2, 1, 0)
            np.testing.assert_allclose(w32, want, rtol=2e-6, atol=1e-9)
        else:
            # ConvTranspose1d [Cin,Cout,K]: (tap j, phase p) <-> kk = u*m + p + pad, m = -off
            u, pad = L.up, L.padding
            total = 0.0
            for j in range(L.num_taps):
                for p in range(u):
                    blk = w32[j][:, p * L.cout:(p + 1) * L.cout]
                    kk = u * (-L.tap_off[j]) + p + pad
                    if (L.tap_mask[j] >> p) & 1:
                        assert 0 <= kk < L.kernel_size
                        np.testing.assert_allclose(blk, w[:, :, kk], rtol=2e-6, atol=1e-9)
                        total += 1
                    else:
                        assert not (0 <= kk < L.kernel_size)
                        assert not blk.any()
            assert total == L.kernel_size   # every kernel element lands exactly once
        if L.w16_off >= 0:
            w16 = packed[L.w16_off: L.w16_off + 2 * L.num_taps * L.n_total * L.cin].view(np.uint16)
            back = (w16.astype(np.uint32) << 16).view(np.float32).reshape(
                L.num_taps, L.n_total, L.cin)
            np.testing.assert_array_equal(back, plan_interp.bf16_round(w32.transpose(0, 2, 1)))


@pytest.mark.parametrize("name", ["rb2_768_stress", "rb1_768_stress", "rb2_1024_warm",
                                  "rb2_768_stress_tiny", "rb1_1024_stress", "rb2_1024_stress"])
def test_fp32_plan_reproduces_reference(name):
    """Layer table + folding + packing + op wiring, executed in numpy fp32: max-abs <= 1e-5
    against the reference output (the device FP32 plan's bar is 1e-4)."""
    g = load_golden(name)
    gen = dropin_in_state(g)
    eng = gen.engine()
    packed = eng.pack_host().numpy()
    plan = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "fp32")
    assert all(op.engine == native.ENGINE_SIMT for op in plan.ops if op.kind == native.OP_TAPGEMM)
    y = plan_interp.run(plan, packed, g.x, g.spk, g.noise)
    assert y.shape == g.y.shape
    assert np.abs(y - g.y).max() < 1e-5


def test_fp32_plan_fresh_state_pre_tanh():
    g = load_golden("rb2_768_fresh")
    gen = dropin_in_state(g)
    eng = gen.engine()
    plan = plan_interp.HostPlan(eng, 1, g.x.shape[2], "fp32")
    y = plan_interp.run(plan, eng.pack_host().numpy(), g.x, g.spk, g.noise)
    assert np.abs(y - g.y).max() < 1e-3      # saturated +-1: sign agreement everywhere


def test_fused_and_unfused_plans_agree(monkeypatch):
    """The fused-stage op (conv_sum trick, on-chip intermediates) is the same arithmetic as the
    six separate tap-GEMMs up to the rounding of the summed conv1 weights."""
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g)
    eng = gen.engine()
    packed = eng.pack_host().numpy()
    fused = plan_interp.run(plan_interp.HostPlan(eng, 2, 24, "bf16"), packed, g.x, g.spk, g.noise)
    monkeypatch.setenv("V2W_NO_FUSE", "1")
    plain = plan_interp.run(plan_interp.HostPlan(eng, 2, 24, "bf16"), packed, g.x, g.spk, g.noise)
    snr = fixtures.snr_db(plain, fused)
    print(f"fused vs unfused (emulated): {snr:.1f} dB")
    assert snr > 60.0


def test_resblock1_pair_fusion(monkeypatch):
    """ResBlock1: every (dilated conv, conv) pair of the C <= 128 stages is one PAIR_FUSED op.  It is
    the same arithmetic with the same rounding points as the two tap-GEMMs it replaces (the
    intermediate is rounded to 16 bits once either way), so the emulated outputs agree to fp32
    summation-order noise; the plan has 36 launches fewer and the same FLOP total."""
    g = load_golden("rb1_768_stress")
    gen = dropin_in_state(g)
    eng = gen.engine()
    packed = eng.pack_host(native.HALF_FP16).numpy()
    B, T = g.x.shape[0], g.x.shape[2]
    fplan = plan_interp.HostPlan(eng, B, T, "fp16")
    pairs = [op for op in fplan.ops if op.kind == native.OP_PAIR_FUSED]
    assert len(pairs) == 36 and sorted({op.stage for op in pairs}) == [1, 2, 3, 4]
    assert len(fplan.ops) == 99 - 36
    assert eng._lib.v2w_model_num_pairs(eng._model) == 36
    native.check(eng._lib.v2w_model_get_pair(eng._model, pairs[0].pair, C.byref(F := native.FusedDesc())))
    # first pair: stage 1 (C = 128) runs the plain program of kernels_fused.cuh, k = 3, dilation 1
    assert (F.ok, F.pair, F.stacked, F.channels, F.num_taps, F.hc, F.h1, F.rc) == (1, 1, 0, 128, 6, 1, 1, 254)
    assert [F.shift[t] for t in range(6)] == [0, 1, 2, 0, 1, 2] and list(F.seg_flags[:2]) == [1, 2]
    F = native.FusedDesc()
    native.check(eng._lib.v2w_model_get_pair(eng._model, pairs[-1].pair, C.byref(F)))
    # last pair of the model: stage 4 (C = 16), k = 11, dilation 5 then 1
    assert (F.ok, F.pair, F.stacked, F.channels, F.num_branches, F.num_taps, F.num_segments) == (1, 1, 1, 16, 1, 22, 2)
    assert (F.hc, F.h1, F.rc, F.a0_boxes, F.t_bufs, F.w_rows) == (5, 25, 246, 5, 2, 22 * 16)
    assert [F.offset[t] for t in range(11)] == [5 * (t - 5) for t in range(11)]
    assert [F.shift[t] for t in range(11, 22)] == list(range(-5, 6)) and list(F.seg_acc[:2]) == [0, 0]
    fused = plan_interp.run(fplan, packed, g.x, g.spk, g.noise)
    monkeypatch.setenv("V2W_NO_FUSE", "1")
    uplan = plan_interp.HostPlan(eng, B, T, "fp16")
    assert len(uplan.ops) == 99 and sum(op.flops for op in uplan.ops) == sum(op.flops for op in fplan.ops)
    plain = plan_interp.run(uplan, packed, g.x, g.spk, g.noise)
    snr = fixtures.snr_db(plain, fused)
    print(f"ResBlock1 pair-fused vs unfused (emulated): {snr:.1f} dB; vs reference {fixtures.snr_db(g.y, fused):.1f} dB")
    assert snr > 100.0
    assert fixtures.snr_db(g.y, fused) >= 45.0


@pytest.mark.parametrize("name,min_snr", [("rb2_768_stress", 45.0), ("rb2_1024_warm", 45.0), ("rb2_1024_stress", 45.0)])
def test_bf16_plan_predicted_snr(name, min_snr):
    """Same op list with the device's bf16 rounding points (bf16 operands, fp32 accumulate,
    fp32 residual stream): the waveform SNR bar of the bf16 path, predicted on the CPU."""
    g = load_golden(name)
    gen = dropin_in_state(g)
    eng = gen.engine()
    plan = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "bf16")
    y = plan_interp.run(plan, eng.pack_host().numpy(), g.x, g.spk, g.noise)
    snr = fixtures.snr_db(g.y, y)
    print(f"{name}: predicted bf16 SNR {snr:.1f} dB")
    assert snr >= min_snr


@pytest.mark.parametrize("name", ["rb2_768_stress", "rb1_768_stress", "rb2_1024_stress"])
def test_fp32tc_plan_predicted_error(name):
    """fp32 values on the tensor cores (V2W_PREC_FP32_TC): every operand split into fp16 hi + lo, three-term
    products, one tap-GEMM per convolution -- emulated on the host with the device's rounding points, the plan
    meets the fp32 bar (max-abs <= 1e-4) with two orders of magnitude to spare."""
    g = load_golden(name)
    gen = dropin_in_state(g)
    eng = gen.engine()
    plan = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "fp32tc")
    kinds = {op.kind for op in plan.ops}
    assert kinds == {native.OP_TRANSPOSE_IN, native.OP_COND_AFFINE, native.OP_TAPGEMM, native.OP_CONV_POST}
    for op in plan.ops:
        if op.kind == native.OP_TAPGEMM:
            L = plan.layers[op.layer]
            assert op.split == 1 and op.engine == native.ENGINE_TC and op.in_pitch == 2 * L.cin
    packed = eng.pack_host(native.HALF_FP16_SPLIT).numpy()
    y = plan_interp.run(plan, packed, g.x, g.spk, g.noise)
    err = np.abs(y - g.y).max()
    print(f"{name}: predicted fp32tc max-abs {err:.2e}")
    assert err < 2e-6
    # the lo halves: fp16(w - fp16(w)) right after each layer's operands; hi + lo holds w to 2^-22
    L = plan.layers[1]
    n = L.num_taps * L.n_total * L.cin
    hi = packed[L.w16_off: L.w16_off + 2 * n].view(np.float16).astype(np.float32)
    lo = packed[L.w16_off + 2 * n: L.w16_off + 4 * n].view(np.float16).astype(np.float32)
    w32 = packed[L.w32_off: L.w32_off + 4 * n].view(np.float32).reshape(L.num_taps, L.cin, L.n_total)
    w = np.ascontiguousarray(w32.transpose(0, 2, 1)).ravel()
    assert np.array_equal(hi, w.astype(np.float16).astype(np.float32))
    assert np.array_equal(lo, (w - hi).astype(np.float16).astype(np.float32))
    assert np.abs(hi + lo - w).max() <= 2.0 ** -21 * np.abs(w).max()
    # a plain fp16 buffer leaves the second half zero
    plain = eng.pack_host(native.HALF_FP16).numpy()
    assert not plain[L.w16_off + 2 * n: L.w16_off + 4 * n].any()


def test_resblock1_row_blocked_stream_wiring(monkeypatch):
    """ResBlock1, fp16 plan: the fp32 stream between the conv pairs is row-blocked -- every pair of the C <= 128 stages
    writes x' blocked and reads a blocked residual from the second pair of a block on; the running sum is blocked
    except in the last stage, whose sum is conv_post's row-major input; the transposed convs in front of stages 1-4
    write X0 blocked, so the first pair of every block reads it blocked as well."""
    g = load_golden("rb1_768_stress")
    eng = dropin_in_state(g).engine()
    plan = plan_interp.HostPlan(eng, 2, 24, "fp16")
    pairs = [op for op in plan.ops if op.kind == native.OP_PAIR_FUSED]
    assert len(pairs) == 36
    for st in (1, 2, 3, 4):
        ops = [op for op in pairs if op.stage == st]
        assert len(ops) == 9
        last = st == 4
        for n, op in enumerate(ops):
            s_in_block, block = n % 3, n // 3
            tail = s_in_block == 2
            assert op.f32_blocked == 1                        # X0 (first pair) or the previous pair's output
            want_out = 0 if (tail and (last or block == 2)) else 1       # last block's tail: 16-bit output only
            assert op.out_blocked == want_out, (st, n)
            assert op.acc_blocked == (1 if (tail and block > 0 and not last) else 0), (st, n)
    ups = [op for op in plan.ops if op.kind == native.OP_TAPGEMM and plan.layers[op.layer].role == 1]
    assert [op.f32_blocked for op in ups] == [0, 1, 1, 1, 1]
    assert [(op.n_tile, op.n_tiles) for op in ups][2] == (256, 1)      # ups[2]: ONE column tile holding its 4 phases
    monkeypatch.setenv("V2W_RB1_NO_BLOCKED", "1")
    monkeypatch.setenv("V2W_BLOCKED_X0_MASK", "0")
    plain = plan_interp.HostPlan(eng, 2, 24, "fp16")
    assert not any(op.f32_blocked or op.out_blocked or op.acc_blocked for op in plain.ops)
    # same arithmetic either way: the layout is invisible to the numpy interpreter
    packed = eng.pack_host(native.HALF_FP16).numpy()
    assert np.array_equal(plan_interp.run(plan, packed, g.x, g.spk, g.noise), plan_interp.run(plain, packed, g.x, g.spk, g.noise))


@pytest.mark.parametrize("name", ["rb1_768_stress", "rb1_1024_stress"])
def test_resblock1_bf16_plan_predicted_snr(name, monkeypatch):
    """ResBlock1 with bf16 operands: rounding the WEIGHTS once costs its 97-conv chain the 45 dB bar.  Default plan:
    operands and weights as bf16 PAIRS (three MMAs per product); $V2W_RB1_BF16=weights: hi + lo weights only (two
    MMAs, activations single bf16 values) -- clears the bar on the stress fixture with 1.5 dB to spare."""
    g = load_golden(name)
    gen = dropin_in_state(g)
    eng = gen.engine()
    full = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "bf16")
    assert full.packed_format == native.HALF_BF16_SPLIT
    gemms = [op for op in full.ops if op.kind == native.OP_TAPGEMM]
    assert len(gemms) == 96 and all(op.split == 1 and op.in_pitch == 2 * full.layers[op.layer].cin for op in gemms)
    snr_full = fixtures.snr_db(g.y, plan_interp.run(full, eng.pack_host(native.HALF_BF16_SPLIT).numpy(), g.x, g.spk, g.noise))
    print(f"{name}: predicted bf16 SNR {snr_full:.1f} dB with bf16 pairs")
    assert snr_full >= 70.0
    monkeypatch.setenv("V2W_RB1_BF16", "weights")
    plan = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "bf16")
    assert plan.packed_format == native.HALF_BF16_SPLIT
    gemms = [op for op in plan.ops if op.kind == native.OP_TAPGEMM]
    assert len(gemms) == 96 and all(op.split == 2 for op in gemms)
    assert not any(op.kind == native.OP_PAIR_FUSED for op in plan.ops)
    y = plan_interp.run(plan, eng.pack_host(native.HALF_BF16_SPLIT).numpy(), g.x, g.spk, g.noise)
    snr = fixtures.snr_db(g.y, y)
    monkeypatch.delenv("V2W_RB1_BF16")
    monkeypatch.setenv("V2W_RB1_BF16_FUSED", "1")
    old = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "bf16")
    assert old.packed_format == native.HALF_BF16 and sum(op.kind == native.OP_PAIR_FUSED for op in old.ops) == 36
    snr_old = fixtures.snr_db(g.y, plan_interp.run(old, eng.pack_host().numpy(), g.x, g.spk, g.noise))
    print(f"{name}: predicted bf16 SNR {snr:.1f} dB with hi + lo weights, {snr_old:.1f} dB single-rounded")
    assert snr >= 45.0 and snr > snr_old
    # the packed lo halves: bf16(w - bf16(w)) right after each layer's operands
    packed = eng.pack_host(native.HALF_BF16_SPLIT).numpy()
    L = plan.layers[3]
    cnt = L.num_taps * L.n_total * L.cin
    as_f32 = lambda off: (packed[off: off + 2 * cnt].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    hi, lo = as_f32(L.w16_off), as_f32(L.w16_off + 2 * cnt)
    w = np.ascontiguousarray(packed[L.w32_off: L.w32_off + 4 * cnt].view(np.float32).reshape(
        L.num_taps, L.cin, L.n_total).transpose(0, 2, 1)).ravel()
    assert np.array_equal(hi, plan_interp.bf16_round(w)) and np.array_equal(lo, plan_interp.bf16_round(w - hi))
    assert np.abs(hi + lo - w).max() <= 2.0 ** -15 * np.abs(w).max()
    # ResBlock2 plans are untouched
    g2 = load_golden("rb2_768_stress_tiny")
    p2 = plan_interp.HostPlan(dropin_in_state(g2).engine(), 1, 8, "bf16")
    assert p2.packed_format == native.HALF_BF16 and not any(op.split for op in p2.ops)


@pytest.mark.parametrize("name,min_snr", [("rb1_768_stress", 55.0), ("rb1_1024_stress", 55.0),
                                          ("rb2_768_stress", 60.0)])
def test_fp16_plan_predicted_snr(name, min_snr):
    """fp16 operands (same tcgen05 rate, 3 more mantissa bits): ResBlock1, which misses the 45 dB
    bar with bf16 operands (40.8 dB on rb1_768_stress), clears it with a wide margin."""
    g = load_golden(name)
    gen = dropin_in_state(g)
    eng = gen.engine()
    plan = plan_interp.HostPlan(eng, g.x.shape[0], g.x.shape[2], "fp16")
    y = plan_interp.run(plan, eng.pack_host(native.HALF_FP16).numpy(), g.x, g.spk, g.noise)
    snr = fixtures.snr_db(g.y, y)
    print(f"{name}: predicted fp16 SNR {snr:.1f} dB")
    assert snr >= min_snr


def test_pack_fp16_bits_and_header():
    g = load_golden("rb2_768_stress_tiny")
    gen = dropin_in_state(g)
    eng = gen.engine()
    pb = eng.pack_host(native.HALF_BF16).numpy()
    ph = eng.pack_host(native.HALF_FP16).numpy()
    assert list(pb[:12].view(np.uint32)) == [0x50573256, 1, 0]
    assert list(ph[:12].view(np.uint32)) == [0x50573256, 1, 1]
    for L in eng.layers():
        if L.w16_off < 0:
            continue
        n = L.num_taps * L.n_total * L.cin
        w32 = ph[L.w32_off: L.w32_off + 4 * n].view(np.float32).reshape(L.num_taps, L.cin, L.n_total)
        got = ph[L.w16_off: L.w16_off + 2 * n].view(np.float16).reshape(L.num_taps, L.n_total, L.cin)
        np.testing.assert_array_equal(got, w32.transpose(0, 2, 1).astype(np.float16))
    # the C-side converter also handles subnormals / overflow like numpy (checked through a layer's
    # weights scaled into those ranges is overkill; the edge cases are covered by conv_post's tiny taps)


def test_default_precision_is_auto():
    assert Generator(hparams(resblock=1)).precision == "bf16"
    assert Generator(hparams(resblock='1')).precision == "fp16"
    gen = Generator(hparams(resblock='1'))
    gen.precision = "fp32"
    assert gen.precision == "fp32"
    with pytest.raises(ValueError):
        gen.precision = "int8"


def test_plan_structure_bf16_rb2(monkeypatch):
    gen = Generator(hparams(resblock=1, num_wv_feat=768))
    eng = gen.engine()
    monkeypatch.setenv("V2W_NO_FUSE", "1")
    plan = plan_interp.HostPlan(eng, 16, 200, "bf16")
    kinds = [op.kind for op in plan.ops]
    assert kinds[0] == native.OP_TRANSPOSE_IN and kinds[1] == native.OP_COND_AFFINE
    assert kinds[-1] == native.OP_CONV_POST
    gemms = [op for op in plan.ops if op.kind == native.OP_TAPGEMM]
    assert len(gemms) == 36 and all(op.engine == native.ENGINE_TC for op in gemms)
    total = sum(op.flops for op in plan.ops)
    assert total == eng.flops_per_frame() * 16 * 200
    # rows double-check: stage outputs 1000, 4000, 16000, 32000, 64000
    assert sorted({op.rows for op in gemms}) == [200, 1000, 4000, 16000, 32000, 64000]
    # default plan: the three C <= 64 stages collapse into one fused op each, same FLOP total
    monkeypatch.delenv("V2W_NO_FUSE")
    fplan = plan_interp.HostPlan(eng, 16, 200, "bf16")
    fused = [op for op in fplan.ops if op.kind == native.OP_MRF_FUSED]
    assert [op.stage for op in fused] == [1, 2, 3, 4]
    assert [op.n_tile for op in fused] == [128, 64, 32, 16]
    # stage 0 (C = 256, no fused kernel): two stacked tap-GEMM launches instead of six (v2w_wide_desc)
    assert len([op for op in fplan.ops if op.kind == native.OP_TAPGEMM]) == 6
    wide = [op for op in fplan.ops if op.kind in (native.OP_MRF_WIDE1, native.OP_MRF_WIDE2)]
    assert [(op.kind, op.stage, op.n_tile, op.n_tiles, op.in_pitch) for op in wide] == \
        [(native.OP_MRF_WIDE1, 0, 256, 3, 1024), (native.OP_MRF_WIDE2, 0, 256, 1, 1024)]
    ups0 = fplan.ops[3]
    assert ups0.kind == native.OP_TAPGEMM and ups0.out_bf16_pitch == 1024
    assert ups0.out_bf16 == wide[0].in_act == wide[1].in_act + 3 * 256 * 2 and wide[0].out_bf16 == wide[1].in_act
    Wd = native.WideDesc()
    native.check(eng._lib.v2w_model_get_wide(eng._model, 0, C.byref(Wd)))
    assert (Wd.ok, Wd.channels, Wd.num_branches, Wd.sum_taps) == (1, 256, 3, 11)
    assert list(Wd.sum_off[:11]) == list(range(-5, 6))
    assert sum(op.flops for op in fplan.ops) == total
    # conv_post + tanh run inside the last stage's final pass: 14 launches, no CONV_POST op
    assert len(fplan.ops) == 14 and fplan.ops[-1].kind == native.OP_MRF_FUSED
    # $V2W_WIDE_MASK=0: the six separate tap-GEMMs of stage 0 again
    monkeypatch.setenv("V2W_WIDE_MASK", "0")
    oplan = plan_interp.HostPlan(eng, 16, 200, "bf16")
    assert len(oplan.ops) == 18 and sum(op.flops for op in oplan.ops) == total
    monkeypatch.delenv("V2W_WIDE_MASK")
    assert [op.post for op in fused] == [0, 0, 0, 1] and fused[-1].out_f32 < 0 and fused[-1].out_bf16 < 0
    F4 = native.FusedDesc()
    native.check(eng._lib.v2w_model_get_fused(eng._model, 4, C.byref(F4)))
    assert (F4.post, F4.hc, F4.rc, F4.channels) == (1, 18, 220, 16)      # 3 more halo rows for the 7-tap conv_post
    F = native.FusedDesc()
    native.check(eng._lib.v2w_model_get_fused(eng._model, 2, C.byref(F)))
    # C = 64: stacked program, one accumulator set, `sum` of the first offset group split off
    T1, T2, A0, A1 = (lambda j: 1 << j), 1 << 4, 1 << 5, (lambda j: 1 << (8 + j))
    assert (F.ok, F.stacked, F.hc, F.h1, F.rc, F.a0_boxes, F.num_taps, F.num_segments) == (1, 1, 15, 5, 226, 5, 35, 7)
    assert (F.t_bufs, F.a0_bufs, list(F.order[:3]), F.w_rows) == (1, 1, [0, 1, 2], 53 * 64)
    assert list(F.seg_type[:7]) == [0, 1, 0, 0, 2, 2, 2] and list(F.seg_ntaps[:7]) == [3, 3, 4, 4, 3, 7, 11]
    assert list(F.seg_dcol[:7]) == [64, 0, 0, 0, 0, 0, 0] and list(F.seg_n[:7]) == [192, 64, 192, 128, 64, 64, 64]
    # T1 is pre-loaded with x + b1 by the epilogue: conv1 accumulates from its first MMA and waits for
    # "T1 pre-loaded" (bit T1(0)); the split-off `sum` taps still overwrite T2
    assert F.preload == 1 and list(F.seg_acc[:7]) == [1, 0, 1, 1, 1, 1, 1]
    assert list(F.seg_wait[:7]) == [T1(0) | A0, T2, 0, 0, A1(0), A1(1), A1(2)]
    assert list(F.seg_commit[:7]) == [T1(0), 0, T1(1), T1(2) | A0, 0, 0, T2]
    assert [F.offset[t] for t in range(14)] == [-1, 0, 1, -1, 0, 1, -3, -2, 2, 3, -5, -4, 4, 5]
    assert [F.shift[t] for t in range(14)] == [5 + o for o in [-1, 0, 1, -1, 0, 1, -3, -2, 2, 3, -5, -4, 4, 5]]
    assert [F.shift[t] for t in range(14, 17)] == [-3, 0, 3] and F.shift[24] == -15 and F.shift[34] == 15
    # C <= 32: two accumulator sets, `sum` stacked into every group (N = 4C, 3C, 2C)
    native.check(eng._lib.v2w_model_get_fused(eng._model, 3, C.byref(F)))
    assert (F.ok, F.stacked, F.channels, F.t_bufs, F.a0_bufs, F.num_taps, F.num_segments) == (1, 1, 32, 2, 2, 32, 6)
    assert list(F.seg_n[:6]) == [128, 96, 64, 32, 32, 32] and list(F.seg_dcol[:6]) == [0] * 6
    assert list(F.seg_wait[:6]) == [T2 | A0, 0, 0, A1(0), A1(1), A1(2)]
    assert F.w_rows == 53 * 32
    # C = 128 does not stack (N <= 256, TMEM): plain program, single buffered: conv_sum right
    # behind conv1_0, then conv2_j / conv1_{j+1}
    native.check(eng._lib.v2w_model_get_fused(eng._model, 1, C.byref(F)))
    assert F.ok and F.channels == 128 and not F.stacked
    assert (F.hc, F.h1, F.rc, F.a0_boxes, F.num_taps, F.num_segments, F.w_rows) == (15, 5, 226, 5, 53, 7, 53 * 128)
    assert list(F.seg_type[:7]) == [0, 1, 2, 0, 2, 0, 2] and list(F.seg_branch[:7]) == [0, -1, 0, 1, 1, 2, 2]
    assert list(F.seg_flags[:7]) == [0, 0, 0, 0, 0, 1, 2]
    # optional schedule: the `sum` convolution in three pieces, one behind every conv1 (a model-level switch)
    monkeypatch.setenv("V2W_FUSED_SUM_PIECES", "3")
    eng3 = Generator(hparams(resblock=1, num_wv_feat=768)).engine()
    native.check(eng3._lib.v2w_model_get_fused(eng3._model, 1, C.byref(F)))
    assert (F.num_taps, F.num_segments) == (53, 9)
    assert list(F.seg_type[:9]) == [0, 1, 2, 0, 1, 2, 0, 1, 2] and list(F.seg_ntaps[:9]) == [3, 3, 3, 7, 4, 7, 11, 4, 11]
    assert list(F.seg_acc[:9]) == [0, 0, 0, 0, 1, 0, 0, 1, 0]
    assert list(F.seg_flags[:9]) == [0, 0, 0, 0, 0, 0, 0, 1, 2]       # the last `sum` piece releases the input tile
    monkeypatch.delenv("V2W_FUSED_SUM_PIECES")
    # transposed convs: column tiles are whole phase groups; a tile issues 2 taps when its phases
    # share a tap set, 3 when the cost model prefers a wider tile with a zero block
    for op in gemms:
        L = plan.layers[op.layer]
        if L.role == native.ROLE_UPSAMPLE:
            assert op.n_tile % L.cout == 0 or L.cout % op.n_tile == 0
            assert all(op.num_taps[t] in (2, 3) for t in range(op.n_tiles))


def test_native_path_has_no_cpu_fallback():
    g = load_golden("rb2_768_stress_tiny")
    gen = dropin_in_state(g)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU"):
            gen(t(g.x), t(g.spk), t(g.noise))
    if native.load().v2w_device_count() == 0:
        eng = gen.engine()
        plan = plan_interp.HostPlan(eng, 1, 4, "bf16")
        rc = eng._lib.v2w_forward(plan.handle, 1, 1, 1, 1, None, None)
        assert rc == -3      # not bound: state error, never a silent success
        rc = eng._lib.v2w_plan_bind(plan.handle, 256, 256)
        assert rc == -2      # no usable GPU


def test_autograd_delegate_matches_reference_output():
    """train()/autograd mode runs stock PyTorch ops in the reference's order: in eval with grad
    enabled it reproduces the golden bit-for-bit (up to thread-order ulps)."""
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g)
    with pytest.warns(RuntimeWarning):
        y = gen(t(g.x), t(g.spk), t(g.noise))
    assert y.requires_grad
    np.testing.assert_allclose(y.detach().numpy(), g.y, rtol=0, atol=2e-6)


def test_cond_only_repack_equals_full_pack():
    """v2w_model_pack_cond writes exactly the conditioning byte range of v2w_model_pack (what the engine
    refreshes in place after train()-mode forwards moved u / v / the running statistics)."""
    import ctypes as C
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False)
    eng = gen.engine()
    full = eng.pack_host().numpy().copy()
    with torch.no_grad():      # move every conditioning tensor a little
        for cbn in gen.cbns:
            cbn.batch_nrom.running_mean.add_(0.05)
            cbn.batch_nrom.running_var.mul_(1.3)
            cbn.layer.weight_u.mul_(0.9)
    full2 = eng.pack_host().numpy()
    nstage = len(gen.cbns)
    ts = eng._state_tensors()[3 * len(gen.conv_modules()):]
    host = [t.detach().float().contiguous() for t in ts]
    cond = (native.CondParams * nstage)()
    for i in range(nstage):
        for (name, _), tt in zip(native.CondParams._fields_, host[8 * i: 8 * i + 8]):
            setattr(cond[i], name, tt.data_ptr())
    off, nb = C.c_int64(0), C.c_int64(0)
    native.check(eng._lib.v2w_model_cond_region(eng._model, C.byref(off), C.byref(nb)))
    region = torch.empty(nb.value, dtype=torch.uint8)
    native.check(eng._lib.v2w_model_pack_cond(eng._model, cond, nstage, region.data_ptr()))
    lo, hi = off.value, off.value + nb.value
    assert np.array_equal(full2[lo:hi], region.numpy())
    assert not np.array_equal(full[lo:hi], full2[lo:hi])
    assert np.array_equal(full[:lo], full2[:lo]) and np.array_equal(full[hi:], full2[hi:])   # nothing else depends on them


def test_train_and_mel_entry_points_validate_before_touching_the_gpu():
    """Error behaviour of the widened C-ABI (v2w_forward_train, v2w_mel_*) that needs no device:
    argument validation first, call-order violations (not bound) second, never a crash."""
    import ctypes as C
    lib = native.load()
    gen = Generator(hparams())
    eng = gen.engine()
    plan = C.c_void_p()
    native.check(lib.v2w_plan_create(eng._model, 2, 8, native.PREC_FP32, C.byref(plan)))
    try:
        st = native.TrainState()
        st.num_stages = 4                      # the model has 5
        st.momentum = 0.1
        one = C.c_void_p(256)                  # any non-null value: never dereferenced on these paths
        rc = lib.v2w_forward_train(plan, one, one, one, one, C.byref(st), None, None)
        assert rc == -1 and "stages" in native.last_error()
        st.num_stages = 5
        rc = lib.v2w_forward_train(plan, one, one, one, one, C.byref(st), None, None)
        assert rc == -1 and "null pointer in stage 0" in native.last_error()
        for i in range(5):
            for name, _ in native.TrainStage._fields_:
                setattr(st.stage[i], name, 256)
        st.momentum = 1.5
        rc = lib.v2w_forward_train(plan, one, one, one, one, C.byref(st), None, None)
        assert rc == -1 and "momentum" in native.last_error()
        st.momentum = 0.1
        rc = lib.v2w_forward_train(plan, one, one, one, one, C.byref(st), None, None)
        assert rc == -3 and "not bound" in native.last_error()
        assert lib.v2w_forward_train(plan, one, one, one, one, None, None, None) == -1
    finally:
        lib.v2w_plan_destroy(plan)
    cfg = native.MelConfig(1024, 256, 1024, 80, 16000, 0.0, 0.0)
    mel = C.c_void_p()
    native.check(lib.v2w_mel_create(C.byref(cfg), 2, 8000, C.byref(mel)))
    try:
        assert lib.v2w_mel_num_frames(mel) == 31 and lib.v2w_mel_workspace_bytes(mel) > 0
        rc = lib.v2w_mel_forward(mel, C.c_void_p(256), C.c_void_p(256), None, None)
        assert rc == -3 and "not bound" in native.last_error()
        assert lib.v2w_mel_bind(mel, C.c_void_p(100)) == -1          # unaligned workspace
    finally:
        lib.v2w_mel_destroy(mel)
    for bad in (native.MelConfig(1024, 250, 1024, 80, 16000, 0.0, 0.0),     # hop % 16
                native.MelConfig(1024, 256, 2048, 80, 16000, 0.0, 0.0),     # win > n_fft
                native.MelConfig(1024, 256, 1024, 80, 16000, 9000.0, 8000.0),   # fmin >= fmax
                native.MelConfig(8192, 16, 8192, 80, 16000, 0.0, 0.0)):     # n_fft / hop > 16 taps
        h = C.c_void_p()
        assert lib.v2w_mel_create(C.byref(bad), 1, 20000, C.byref(h)) == -1 and not h.value
    h = C.c_void_p()
    assert lib.v2w_mel_create(C.byref(cfg), 1, 300, C.byref(h)) == -1       # shorter than the reflect padding


def test_pack_is_independent_of_the_thread_count(monkeypatch):
    """The weight fold runs as parallel (layer, tap) tasks writing disjoint regions: same bytes with 1, 3
    and the default number of host threads, both 16-bit formats."""
    gen = Generator(hparams(resblock='1'))
    eng = gen.engine()
    monkeypatch.setenv("V2W_PACK_THREADS", "1")
    one = [eng.pack_host(f).numpy().copy() for f in (native.HALF_BF16, native.HALF_FP16)]
    monkeypatch.setenv("V2W_PACK_THREADS", "3")
    three = [eng.pack_host(f).numpy().copy() for f in (native.HALF_BF16, native.HALF_FP16)]
    monkeypatch.delenv("V2W_PACK_THREADS")
    dflt = [eng.pack_host(f).numpy() for f in (native.HALF_BF16, native.HALF_FP16)]
    for a, b, c in zip(one, three, dflt):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert not np.array_equal(one[0], one[1])
