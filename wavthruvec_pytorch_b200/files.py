"""Feature files in, WAV files out (the producers and consumers either
side of the generator, in the reference's own on-disk formats).

    <feat_dir>/<speaker>/<utt>.npy     wav2vec 2.0 last_hidden_state, float32 [1, T, C]   (prepare_data.py:75-87)
    <spk_emb_dir>/<speaker>.pth        ECAPA embedding, torch tensor [1, 1, 192]           (vec2wav/pre_spk_emb.py:20-33)
    list file lines                    "<speaker>/<utt>.npy|<text>|<speaker>"              (prepare_data.py:89-92)

The reference consumes these through VocoderDataset / collate_fn_tensor (vec2wav/dataset.py:181-228)
and draws one noise vector per utterance for every batch (vec2wav/train.py:165-166, 259-260; here with a
seeded generator on the device).  `synthesize_files` does the same for
generation, as a three-stage host pipeline around the native variable-length path:

    reader threads:  np.load, speaker embeddings cached per speaker
    main thread:     packs a batch (frame budget) into a pinned arena = the x_rows layout of
                     v2w_forward_ragged, ONE H2D copy, device noise, forward_ragged(pcm16), D2H
    writer thread:   16-bit mono RIFF/WAVE files

Threads overlap because np.load / file writes / CUDA calls release the GIL.  Host plumbing only; the
arithmetic is libv2w's.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import varlen


@dataclass
class Item:
    name: str          # "<speaker>/<utt>" (no extension)
    speaker: str


def read_list(list_path: str) -> List[Item]:
    """Items of a prepare_data.py list file ("spk/utt.npy|text|spk")."""
    items = []
    with open(list_path, encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            parts = line.split("|")
            if len(parts) < 3:
                raise ValueError(f"malformed list line: {line!r}")
            items.append(Item(os.path.splitext(parts[0])[0], parts[-1]))
    return items


def scan_features(feat_dir: str) -> List[Item]:
    """Every <speaker>/<utt>.npy under feat_dir, sorted."""
    items = []
    for spk in sorted(os.listdir(feat_dir)):
        d = os.path.join(feat_dir, spk)
        if os.path.isdir(d):
            items += [Item(f"{spk}/{f[:-4]}", spk) for f in sorted(os.listdir(d)) if f.endswith(".npy")]
    return items


def load_features(path: str, num_wv_feat: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One utterance's features as a [T, C] fp32 tensor (dataset.py:181: np.load + squeeze), read by
    libv2w's .npy reader (v2w_npy_read_f32).  The ctypes call releases the GIL, so reader threads
    overlap.  `out`: optional [>= T, C] fp32 buffer (e.g. pinned) to read into; a view of it is returned."""
    import ctypes as C
    from . import native
    lib = native.load()
    bpath = os.fsencode(path)
    if out is None:
        rows, cols = C.c_int64(0), C.c_int64(0)
        try:
            native.check(lib.v2w_npy_shape(bpath, C.byref(rows), C.byref(cols)))
        except RuntimeError as e:
            raise ValueError(str(e)) from None
        if cols.value != num_wv_feat:
            raise ValueError(f"{path}: {cols.value} feature channels, expected {num_wv_feat}")
        out = torch.empty(rows.value, num_wv_feat, dtype=torch.float32)
    if out.dtype != torch.float32 or out.dim() != 2 or out.shape[1] != num_wv_feat or not out.is_contiguous():
        raise ValueError("out must be a contiguous [rows, num_wv_feat] fp32 tensor")
    n = C.c_int64(0)
    try:
        native.check(lib.v2w_npy_read_f32(bpath, out.data_ptr(), out.shape[0], num_wv_feat, C.byref(n)))
    except RuntimeError as e:
        raise ValueError(str(e)) from None
    return out[:n.value]


class SpeakerCache:
    """<spk_emb_dir>/<speaker>.pth -> [spk_dim] fp32, loaded once per speaker."""

    def __init__(self, spk_emb_dir: str, spk_dim: int):
        self.dir, self.dim = spk_emb_dir, spk_dim
        self._cache: Dict[str, torch.Tensor] = {}
        self._lock = threading.Lock()

    def get(self, speaker: str) -> torch.Tensor:
        with self._lock:
            e = self._cache.get(speaker)
        if e is None:
            e = torch.load(os.path.join(self.dir, speaker + ".pth"), map_location="cpu", weights_only=True)
            e = e.detach().to(torch.float32).reshape(-1)       # dataset.py:183-186: squeeze().squeeze()
            if e.numel() != self.dim:
                raise ValueError(f"speaker {speaker}: embedding has {e.numel()} values, expected {self.dim}")
            with self._lock:
                self._cache[speaker] = e
        return e


@dataclass
class Stats:
    utterances: int = 0
    audio_s: float = 0.0
    wall_s: float = 0.0
    batches: int = 0

    @property
    def audio_s_per_s(self) -> float:
        return self.audio_s / self.wall_s if self.wall_s > 0 else 0.0


def synthesize_files(gen, items: Sequence[Item], feat_dir: str, spk_emb_dir: str, out_dir: str,
                     max_batch_frames: int = 16000, max_batch_utts: int = 128, readers: int = 2,
                     seed: int = 1234, sampling_rate: int = 16000, max_utt_frames: int = 2000) -> Stats:
    """Generate <out_dir>/<speaker>/<utt>.wav for every item.  `gen`: drop-in Generator in eval()
    mode on a CUDA device.  A batch closes at `max_batch_frames` feature frames or
    `max_batch_utts` utterances, whichever comes first (items keep their order)."""
    if gen.training:
        raise RuntimeError("synthesize_files is the eval()-mode generation path")
    dev = next(gen.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("the native Vec2Wav generation path runs on sm_100 GPUs only and has no CPU fallback")
    h = gen.h
    spk_cache = SpeakerCache(spk_emb_dir, h.spk_dim)
    rng = torch.Generator(device=dev)
    rng.manual_seed(seed)
    to_write: "queue.Queue" = queue.Queue(maxsize=8)
    errors: List[BaseException] = []
    made_dirs = set()
    # reader threads take item indices in order; results are re-ordered by index in the main thread
    next_index = [0]
    lock = threading.Lock()

    def reader():
        try:
            while not errors:
                # buffer first, index second: the thread that owns the next index in line always
                # already owns a buffer, so out-of-order arrivals cannot starve it
                buf = scratch.get()                                  # pre-faulted pinned buffer, recycled
                with lock:
                    i = next_index[0]
                    next_index[0] += 1
                if i >= len(items):
                    scratch.put(buf)
                    return
                it = items[i]
                feats = load_features(os.path.join(feat_dir, it.name + ".npy"), h.num_wv_feat, out=buf)
                loaded.put((i, feats, spk_cache.get(it.speaker), buf))   # file -> host buffer without the GIL
        except BaseException as e:   # surfaced in the main thread
            errors.append(e)
            loaded.put((-1, None, None, None))

    def writer():
        try:
            while True:
                job = to_write.get()
                if job is None:
                    return
                done_event, pcm_host, names, lengths, pcm_buf = job
                done_event.synchronize()
                off = 0
                for name, n in zip(names, lengths):
                    path = os.path.join(out_dir, name + ".wav")
                    d = os.path.dirname(path)
                    if d not in made_dirs:
                        os.makedirs(d, exist_ok=True)
                        made_dirs.add(d)
                    varlen.write_wav(path, pcm_host[off:off + n], sampling_rate)
                    off += n
                pcm_pool.put(pcm_buf)
        except BaseException as e:
            errors.append(e)

    # Reader scratch: pinned [max_utt_frames, C] buffers handed reader -> main thread -> reader.  No
    # per-file allocation (a fresh 1 MB tensor per file page-faults and serialises the readers).
    max_utt_frames = min(max_utt_frames, max_batch_frames)
    n_scratch = 4 * max(1, readers) + 2
    scratch: "queue.Queue" = queue.Queue()
    for _ in range(n_scratch):
        scratch.put(torch.empty(max_utt_frames, h.num_wv_feat, dtype=torch.float32).pin_memory())
    loaded = queue.Queue(maxsize=n_scratch)
    threads = [threading.Thread(target=reader, daemon=True) for _ in range(max(1, readers))]
    wt = threading.Thread(target=writer, daemon=True)
    for t in threads + [wt]:
        t.start()

    stats = Stats()
    t0 = time.perf_counter()
    pending: Dict[int, tuple] = {}
    want = 0
    batch: List[tuple] = []
    frames = 0
    eng = gen.engine()
    up = eng.total_upsample()
    # Two pinned arenas [frames, C] that alternate between batches: a batch's rows are packed here --
    # exactly the x_rows layout of v2w_forward_ragged -- and go to the device in ONE copy.
    # Plan shapes are quantised (utterance count to a multiple of 8 by appending 1-frame dummies whose
    # output is dropped, frames to a multiple of 32) so that successive batches reuse plans; a longer
    # plan only means the longest utterance is padded like the shorter ones, which is what the
    # reference does to it in any batch that holds a longer one.
    arena_rows = max_batch_frames + max_batch_utts + 8
    arenas = [torch.empty(arena_rows, h.num_wv_feat, dtype=torch.float32).pin_memory() for _ in range(2)]
    arena_free = [torch.cuda.Event(), torch.cuda.Event()]
    n_flush = [0]
    # pinned PCM buffers cycle main thread -> writer -> main thread
    pcm_pool: "queue.Queue" = queue.Queue()
    for _ in range(3):
        pcm_pool.put(torch.empty(arena_rows * up, dtype=torch.int16).pin_memory())

    # batch under construction: (item index, frames, speaker embedding); its rows are already in the arena
    def add(i, f, s_emb, buf):
        nonlocal frames
        if not batch:
            arena_free[n_flush[0] & 1].synchronize()   # the H2D copy that last read this arena is done
        n = f.shape[0]
        arenas[n_flush[0] & 1][frames:frames + n].copy_(f)
        scratch.put(buf)                               # the reader's buffer goes straight back
        batch.append((i, n, s_emb))
        frames += n

    def flush():
        nonlocal batch, frames
        if not batch:
            return
        a = n_flush[0] & 1
        n_flush[0] += 1
        arena = arenas[a]
        lengths = [n for _, n, _ in batch]
        off = frames
        real = len(batch)
        dummies = (-real) % 8
        arena[off:off + dummies].zero_()
        lengths += [1] * dummies
        rows_dev = arena[:off + dummies].to(dev, non_blocking=True)
        arena_free[a].record()
        spk = torch.zeros(real + dummies, h.spk_dim)
        for j, (_, _, s_emb) in enumerate(batch):
            spk[j] = s_emb
        spk = spk.to(dev)
        noise = torch.randn(real + dummies, h.noise_dim, device=dev, generator=rng)
        plan_frames = -(-max(lengths) // 32) * 32
        packed = eng.forward_ragged(rows_dev, lengths, spk, noise, plan_frames, pcm16=True)
        n_valid = off * up
        pcm_buf = pcm_pool.get()
        pcm_host = pcm_buf[:n_valid]
        pcm_host.copy_(packed[:n_valid], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        to_write.put((ev, pcm_host, [items[i].name for i, _, _ in batch],
                      [n * up for n in lengths[:real]], pcm_buf))
        stats.utterances += real
        stats.audio_s += n_valid / sampling_rate
        stats.batches += 1
        batch, frames = [], 0

    with torch.no_grad():
        while want < len(items):
            while want not in pending:
                try:
                    i, f, s, buf = loaded.get(timeout=120)
                except queue.Empty:
                    raise RuntimeError(f"feature readers made no progress for 120 s (waiting for item {want})") from None
                if errors:
                    raise errors[0]
                pending[i] = (f, s, buf)
            f, s, buf = pending.pop(want)
            if batch and (frames + f.shape[0] > max_batch_frames or len(batch) >= max_batch_utts):
                flush()
            add(want, f, s, buf)
            want += 1
        flush()
    to_write.put(None)
    wt.join()
    if errors:
        raise errors[0]
    torch.cuda.synchronize(dev)
    stats.wall_s = time.perf_counter() - t0
    return stats
