"""Feature files in, WAV files out: file formats on CPU, the whole pipeline on GPU."""
from __future__ import annotations

import struct

import numpy as np
import pytest
import torch

from helpers import dropin_in_state, load_golden
from wavthruvec_pytorch_b200 import files, varlen


def _make_tree(tmp_path, C, lengths, speakers, seed=5):
    rng = np.random.default_rng(seed)
    feat_dir, spk_dir = tmp_path / "w2v_feat", tmp_path / "spk_emb"
    spk_dir.mkdir()
    items, lines = [], []
    for s in sorted(set(speakers)):
        (feat_dir / s).mkdir(parents=True)
        torch.save(torch.from_numpy(rng.standard_normal((1, 1, 192)).astype(np.float32)), spk_dir / f"{s}.pth")
    for i, (n, s) in enumerate(zip(lengths, speakers)):
        name = f"{s}/utt{i:03d}"
        np.save(feat_dir / f"{name}.npy", rng.standard_normal((1, n, C)).astype(np.float32))   # prepare_data.py:87
        items.append(files.Item(name, s))
        lines.append(f"{name}.npy|some text|{s}")
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(lines) + "\n", encoding="utf-8")
    return str(feat_dir), str(spk_dir), str(lst), items


def test_file_formats(tmp_path):
    feat_dir, spk_dir, lst, items = _make_tree(tmp_path, 768, [5, 9, 3], ["SSB0005", "SSB0009", "SSB0005"])
    assert files.read_list(lst) == items
    assert files.scan_features(feat_dir) == sorted(items, key=lambda it: it.name)
    f = files.load_features(f"{feat_dir}/{items[1].name}.npy", 768)
    assert tuple(f.shape) == (9, 768) and f.dtype == torch.float32
    assert np.array_equal(f.numpy(), np.load(f"{feat_dir}/{items[1].name}.npy")[0])
    with pytest.raises(ValueError):
        files.load_features(f"{feat_dir}/{items[1].name}.npy", 1024)
    # reading into a caller-provided buffer; .npy variants the reader must take or refuse
    buf = torch.full((16, 768), 7.0)
    v = files.load_features(f"{feat_dir}/{items[0].name}.npy", 768, out=buf)
    assert tuple(v.shape) == (5, 768) and v.data_ptr() == buf.data_ptr() and float(buf[5, 0]) == 7.0
    with pytest.raises(ValueError, match="do not fit"):
        files.load_features(f"{feat_dir}/{items[1].name}.npy", 768, out=torch.empty(4, 768))
    a = np.arange(6 * 768, dtype=np.float32).reshape(6, 768)
    np.save(tmp_path / "two_d.npy", a)                                  # [T, C] without the batch axis
    assert np.array_equal(files.load_features(str(tmp_path / "two_d.npy"), 768).numpy(), a)
    np.save(tmp_path / "f64.npy", a.astype(np.float64))
    np.save(tmp_path / "fortran.npy", np.asfortranarray(a))
    np.save(tmp_path / "four_d.npy", a.reshape(1, 1, 6, 768))
    (tmp_path / "junk.npy").write_bytes(b"not numpy at all")
    for bad in ("f64", "fortran", "four_d", "junk", "missing"):
        with pytest.raises(ValueError):
            files.load_features(str(tmp_path / f"{bad}.npy"), 768)
    cache = files.SpeakerCache(spk_dir, 192)
    e = cache.get("SSB0005")
    assert tuple(e.shape) == (192,) and cache.get("SSB0005") is e
    with pytest.raises(ValueError):
        files.SpeakerCache(spk_dir, 64).get("SSB0009")
    (tmp_path / "bad.txt").write_text("no separators here\n")
    with pytest.raises(ValueError):
        files.read_list(str(tmp_path / "bad.txt"))


def _read_wav(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    n = struct.unpack("<I", raw[40:44])[0]
    return np.frombuffer(raw[44:44 + n], dtype="<i2")


@pytest.mark.gpu
def test_files_to_wav_matches_generate_varlen(tmp_path):
    _run_files_roundtrip(tmp_path, "cuda:0")


def _run_files_roundtrip(tmp_path, device):
    g = load_golden("rb2_768_stress")
    gen = dropin_in_state(g, cache=False).to(device)
    lengths = [37, 12, 50, 1, 29, 44, 8]
    speakers = ["A", "B", "A", "C", "B", "A", "C"]
    feat_dir, spk_dir, lst, items = _make_tree(tmp_path, gen.h.num_wv_feat, lengths, speakers)
    out_dir = tmp_path / "wav"
    stats = files.synthesize_files(gen, files.read_list(lst), feat_dir, spk_dir, str(out_dir), seed=77, readers=3)
    assert stats.utterances == 7 and stats.batches == 1 and abs(stats.audio_s - sum(lengths) / 50) < 1e-9
    # the same batch through generate_varlen with the same device noise
    rng = torch.Generator(device=device)
    rng.manual_seed(77)
    noise = torch.randn(7, gen.h.noise_dim, device=device, generator=rng)
    cache = files.SpeakerCache(spk_dir, gen.h.spk_dim)
    feats = [files.load_features(f"{feat_dir}/{it.name}.npy", gen.h.num_wv_feat).to(device) for it in items]
    spk = torch.stack([cache.get(it.speaker) for it in items]).to(device)
    spk = torch.cat([spk, torch.zeros(1, gen.h.spk_dim, device=device)])     # 7 real + 1 dummy (count -> 8)
    rng.manual_seed(77)
    noise = torch.randn(8, gen.h.noise_dim, device=device, generator=rng)
    feats.append(torch.zeros(1, gen.h.num_wv_feat, device=device))
    packed = gen.engine().forward_ragged(torch.cat(feats), lengths + [1], spk, noise, plan_frames=64, pcm16=True)
    want, off = [], 0
    for n in lengths:
        want.append(packed[off * 320:(off + n) * 320])
        off += n
    for it, w, n in zip(items, want, lengths):
        got = _read_wav(out_dir / f"{it.name}.wav")
        assert got.size == n * 320 and np.array_equal(got, w.cpu().numpy())
    # small batches (frame budget) produce the same files up to the noise draw order: lengths only
    out2 = tmp_path / "wav2"
    stats2 = files.synthesize_files(gen, items, feat_dir, spk_dir, str(out2), max_batch_frames=60, seed=77)
    assert stats2.batches > 1 and stats2.utterances == 7
    for it, n in zip(items, lengths):
        assert _read_wav(out2 / f"{it.name}.wav").size == n * 320


def test_batch_io(tmp_path):
    """The per-batch library calls of the file pipeline: headers, payloads into one packed row buffer,
    WAV files out of one packed PCM buffer -- against numpy / scipy readers."""
    import ctypes as C
    from scipy.io import wavfile
    from wavthruvec_pytorch_b200 import native
    rng = np.random.default_rng(3)
    lengths = [5, 9, 1, 33, 2]
    paths, arrs = [], []
    for i, n in enumerate(lengths):
        a = rng.standard_normal((1, n, 768)).astype(np.float32)
        np.save(tmp_path / f"u{i}.npy", a)
        paths.append(str(tmp_path / f"u{i}.npy"))
        arrs.append(a[0])
    rows, cols = files.feature_shapes(paths, threads=3)
    assert rows.tolist() == lengths and set(cols.tolist()) == {768}
    lib = native.load()
    i64p = C.POINTER(C.c_int64)
    dst = np.full((sum(lengths) + 2, 768), 7.0, dtype=np.float32)
    off = np.zeros(len(lengths), dtype=np.int64)
    np.cumsum(rows[:-1], out=off[1:])
    native.check(lib.v2w_npy_read_batch(files._c_paths(paths), len(paths), dst.ctypes.data, off.ctypes.data_as(i64p),
                                        rows.ctypes.data_as(i64p), 768, 4))
    assert np.array_equal(dst[:-2], np.concatenate(arrs)) and float(dst[-1, 0]) == 7.0
    wrong = rows.copy()
    wrong[1] += 1                                             # layout computed from a stale header
    with pytest.raises(native.V2WError, match="expected"):
        native.check(lib.v2w_npy_read_batch(files._c_paths(paths), len(paths), dst.ctypes.data, off.ctypes.data_as(i64p),
                                            wrong.ctypes.data_as(i64p), 768, 2))
    with pytest.raises(ValueError, match="cannot open"):
        files.feature_shapes(paths + [str(tmp_path / "missing.npy")], threads=2)
    assert files.feature_shapes([], threads=2)[0].size == 0
    pcm = rng.integers(-32768, 32767, size=1000).astype(np.int16)
    wavs = [str(tmp_path / f"w{i}.wav") for i in range(3)]
    so, sn = np.array([0, 100, 400], dtype=np.int64), np.array([100, 300, 600], dtype=np.int64)
    native.check(lib.v2w_wav_write_batch(files._c_paths(wavs), 3, pcm.ctypes.data, so.ctypes.data_as(i64p),
                                         sn.ctypes.data_as(i64p), 16000, 2))
    for p, o, n in zip(wavs, so, sn):
        sr, w = wavfile.read(p)
        assert sr == 16000 and w.dtype == np.int16 and np.array_equal(w, pcm[o:o + n])
        assert np.array_equal(_read_wav(p), pcm[o:o + n])
    with pytest.raises(native.V2WError, match="cannot create"):
        native.check(lib.v2w_wav_write_batch(files._c_paths([str(tmp_path / "no_dir" / "x.wav")]), 1, pcm.ctypes.data,
                                             so.ctypes.data_as(i64p), sn.ctypes.data_as(i64p), 16000, 1))


def test_plan_batches():
    assert files.plan_batches([], 100, 4) == []
    assert files.plan_batches([10, 20, 30], 100, 8) == [(0, 3)]
    assert files.plan_batches([60, 50, 40, 10, 200, 5], 100, 8) == [(0, 1), (1, 4), (4, 5), (5, 6)]
    assert files.plan_batches([1] * 10, 100, 4) == [(0, 4), (4, 8), (8, 10)]
    rng = np.random.default_rng(0)
    rows = rng.integers(1, 500, size=300).tolist()
    b = files.plan_batches(rows, 4000, 32)
    assert b[0][0] == 0 and b[-1][1] == 300 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    for lo, hi in b:
        assert hi - lo <= 32 and (sum(rows[lo:hi]) <= 4000 or hi - lo == 1)
        if hi < 300:   # greedy: the next utterance did not fit
            assert sum(rows[lo:hi]) + rows[hi] > 4000 or hi - lo == 32


def test_plan_batches_bounds_the_padded_plan_rectangle():
    """127 short utterances must not ride with one 2000-frame utterance in a 128 x 2016 plan."""
    rows = [20] * 127 + [2000]
    unbounded = files.plan_batches(rows, 16000, 128)
    assert unbounded == [(0, 128)]
    bounded = files.plan_batches(rows, 16000, 128, max_plan_frames=32000)
    assert bounded[-1] == (127, 128) and bounded[0][0] == 0
    for lo, hi in bounded:
        utts = -(-(hi - lo) // 8) * 8
        longest = -(-max(rows[lo:hi]) // 32) * 32
        assert utts * longest <= 32000 or hi - lo == 1
    # homogeneous corpora are unaffected
    rows = [200] * 64
    assert files.plan_batches(rows, 6400, 32, max_plan_frames=12800) == files.plan_batches(rows, 6400, 32)


@pytest.mark.gpu
def test_files_pipeline_on_a_non_current_device(tmp_path):
    """the generator may live on a device that is not the caller's current device; the pipeline's
    events must be recorded on that device's stream (wrong-device events let the prefetch thread overwrite a
    pinned arena mid-copy)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    _run_files_roundtrip(tmp_path, "cuda:1")
