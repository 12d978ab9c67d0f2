"""Feature files in, WAV files out (the producers and consumers either
side of the generator, in the reference's own on-disk formats).

    <feat_dir>/<speaker>/<utt>.npy     wav2vec 2.0 last_hidden_state, float32 [1, T, C]   (prepare_data.py:75-87)
    <spk_emb_dir>/<speaker>.pth        ECAPA embedding, torch tensor [1, 1, 192]           (vec2wav/pre_spk_emb.py:20-33)
    list file lines                    "<speaker>/<utt>.npy|<text>|<speaker>"              (prepare_data.py:89-92)

The reference consumes these through VocoderDataset / collate_fn_tensor (vec2wav/dataset.py:181-228)
and draws one noise vector per utterance for every batch (vec2wav/train.py:165-166, 259-260; here with a
seeded generator on the device).  `synthesize_files` does the same for
generation, as a three-stage host pipeline around the native variable-length path, with three library
calls per BATCH and no per-file work in Python:

    prefetch thread: v2w_npy_read_batch -- worker threads inside libv2w read every payload of batch k+1
                     straight to its final offset of a pinned arena (= the x_rows layout of
                     v2w_forward_ragged); speaker embeddings cached per speaker
    main thread:     ONE H2D copy, device noise, forward_ragged(pcm16), ONE D2H copy
    writer thread:   v2w_wav_write_batch -- 16-bit mono RIFF/WAVE files of batch k-1

The ctypes calls release the GIL, so the three stages overlap.  Host plumbing only; the arithmetic is
libv2w's.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

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


def _c_paths(paths: Sequence[str]):
    import ctypes as C
    return (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])


def feature_shapes(paths: Sequence[str], threads: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    """(frames, channels) of every feature file, read by libv2w worker threads in ONE call
    (v2w_npy_shapes)."""
    import ctypes as C
    from . import native
    n = len(paths)
    rows, cols = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    if n:
        try:
            native.check(native.load().v2w_npy_shapes(
                _c_paths(paths), n, rows.ctypes.data_as(C.POINTER(C.c_int64)),
                cols.ctypes.data_as(C.POINTER(C.c_int64)), int(threads)))
        except RuntimeError as e:
            raise ValueError(str(e)) from None
    return rows, cols


def plan_batches(rows: Sequence[int], max_batch_frames: int, max_batch_utts: int,
                 max_plan_frames: Optional[int] = None) -> List[Tuple[int, int]]:
    """Item ranges [lo, hi) in order: a batch closes when the next utterance would exceed the frame
    budget or the utterance cap (a single over-long utterance still gets a batch of its own), or -- with
    `max_plan_frames` -- when the padded rectangle of its plan (utterances rounded up to 8 x longest
    utterance rounded up to 32: what the workspace is sized by) would."""
    batches: List[Tuple[int, int]] = []
    lo, frames, longest = 0, 0, 0
    for i, n in enumerate(rows):
        n = int(n)
        if i > lo:
            utts = -(-(i - lo + 1) // 8) * 8
            rect = utts * (-(-max(longest, n) // 32) * 32)
            if (frames + n > max_batch_frames or i - lo >= max_batch_utts
                    or (max_plan_frames is not None and rect > max_plan_frames)):
                batches.append((lo, i))
                lo, frames, longest = i, 0, 0
        frames += n
        longest = max(longest, n)
    if len(rows):
        batches.append((lo, len(rows)))
    return batches


def synthesize_files(gen, items: Sequence[Item], feat_dir: str, spk_emb_dir: str, out_dir: str,
                     max_batch_frames: int = 16000, max_batch_utts: int = 128, readers: int = 4,
                     seed: int = 1234, sampling_rate: int = 16000, max_utt_frames: int = 2000,
                     reader_timeout_s: Optional[float] = None, max_plan_frames: Optional[int] = None) -> Stats:
    """Generate <out_dir>/<speaker>/<utt>.wav for every item.  `gen`: drop-in Generator in eval()
    mode on a CUDA device.  A batch closes at `max_batch_frames` feature frames or
    `max_batch_utts` utterances, whichever comes first (items keep their order).

    Per BATCH the host does three library calls and no per-file work: v2w_npy_read_batch (`readers`
    worker threads, every payload lands at its final offset of a pinned row arena = the x_rows layout
    of v2w_forward_ragged), the ragged forward with PCM16 output, v2w_wav_write_batch.  A prefetch
    thread fills the arena of batch k+1 and a writer thread emits batch k-1 while batch k is on the GPU.

    The workspace of a batch is sized utterances x longest utterance (the plan is a padded rectangle), so a
    batch also closes when that product would exceed `max_plan_frames` (default 2 x max_batch_frames): 127
    short utterances can no longer ride with one 2000-frame utterance in a 128 x 2016 plan.
    `reader_timeout_s`: give up when the reader delivers nothing for this long (default: wait as long as
    the reader thread is alive)."""
    import ctypes as C
    from . import native
    if gen.training:
        raise RuntimeError("synthesize_files is the eval()-mode generation path")
    dev = next(gen.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("the native Vec2Wav generation path runs on sm_100 GPUs only and has no CPU fallback")
    lib = native.load()
    h = gen.h
    stats = Stats()
    t0 = time.perf_counter()
    n_items = len(items)
    if n_items == 0:
        return stats
    readers = max(1, int(readers))
    feat_paths = [os.path.join(feat_dir, it.name + ".npy") for it in items]
    wav_paths = [os.path.join(out_dir, it.name + ".wav") for it in items]
    rows, cols = feature_shapes(feat_paths, readers)
    max_utt_frames = min(max_utt_frames, max_batch_frames)
    for i in range(n_items):
        if cols[i] != h.num_wv_feat:
            raise ValueError(f"{feat_paths[i]}: {cols[i]} feature channels, expected {h.num_wv_feat}")
        if rows[i] > max_utt_frames:
            raise ValueError(f"{feat_paths[i]}: {rows[i]} frames do not fit the {max_utt_frames}-frame limit")
    batches = plan_batches(rows, max_batch_frames, max_batch_utts,
                           max_plan_frames if max_plan_frames is not None else 2 * max_batch_frames)
    for d in sorted({os.path.dirname(p) for p in wav_paths}):
        os.makedirs(d, exist_ok=True)

    spk_cache = SpeakerCache(spk_emb_dir, h.spk_dim)
    rng = torch.Generator(device=dev)
    rng.manual_seed(seed)
    eng = gen.engine()
    up = eng.total_upsample()
    errors: List[BaseException] = []
    # Two pinned arenas [frames, C] that alternate between batches.  Plan shapes are quantised
    # (utterance count to a multiple of 8 by appending 1-frame dummies whose output is dropped, frames
    # to a multiple of 32) so that successive batches reuse plans; a longer plan only means the longest
    # utterance is padded like the shorter ones, which is what the reference does to it in any batch
    # that holds a longer one.
    arena_rows = max_batch_frames + max_batch_utts + 8
    arenas = [torch.empty(arena_rows, h.num_wv_feat, dtype=torch.float32).pin_memory() for _ in range(2)]
    arena_tokens = [queue.Queue(), queue.Queue()]     # event of the H2D copy that last read the arena
    for q in arena_tokens:
        q.put(None)
    ready: "queue.Queue" = queue.Queue(maxsize=1)
    to_write: "queue.Queue" = queue.Queue(maxsize=4)
    pcm_pool: "queue.Queue" = queue.Queue()           # pinned PCM buffers cycle main -> writer -> main
    for _ in range(3):
        pcm_pool.put(torch.empty(arena_rows * up, dtype=torch.int16).pin_memory())
    i64p = C.POINTER(C.c_int64)

    def prefetch():
        try:
            for k, (lo, hi) in enumerate(batches):
                a = k & 1
                ev = arena_tokens[a].get()
                if ev is not None:
                    ev.synchronize()                  # the H2D copy of batch k-2 has left this arena
                if errors:
                    return
                r = np.ascontiguousarray(rows[lo:hi])
                off = np.zeros(hi - lo, dtype=np.int64)
                np.cumsum(r[:-1], out=off[1:])
                try:
                    native.check(lib.v2w_npy_read_batch(
                        _c_paths(feat_paths[lo:hi]), hi - lo, arenas[a].data_ptr(), off.ctypes.data_as(i64p),
                        r.ctypes.data_as(i64p), h.num_wv_feat, readers))
                except RuntimeError as e:
                    raise ValueError(str(e)) from None
                real = hi - lo
                spk = torch.zeros(real + (-real) % 8, h.spk_dim)
                for j in range(real):
                    spk[j] = spk_cache.get(items[lo + j].speaker)
                ready.put((k, spk))
        except BaseException as e:   # surfaced in the main thread
            errors.append(e)
            ready.put(None)

    def writer():
        try:
            while True:
                job = to_write.get()
                if job is None:
                    return
                done_event, pcm_buf, lo, hi = job
                done_event.synchronize()
                samples = np.ascontiguousarray(rows[lo:hi]) * up
                off = np.zeros(hi - lo, dtype=np.int64)
                np.cumsum(samples[:-1], out=off[1:])
                try:
                    native.check(lib.v2w_wav_write_batch(
                        _c_paths(wav_paths[lo:hi]), hi - lo, pcm_buf.data_ptr(), off.ctypes.data_as(i64p),
                        samples.ctypes.data_as(i64p), int(sampling_rate), max(2, readers // 2)))
                finally:
                    pcm_pool.put(pcm_buf)
        except BaseException as e:
            errors.append(e)

    pt = threading.Thread(target=prefetch, daemon=True)
    wt = threading.Thread(target=writer, daemon=True)
    pt.start()
    wt.start()
    failure: Optional[BaseException] = None
    try:
        # everything below runs on `dev` and on ITS current stream, whatever the caller's current device is
        # (events recorded on another device's stream would not cover these copies)
        with torch.cuda.device(dev), torch.no_grad():
            stream = torch.cuda.current_stream(dev)
            for k, (lo, hi) in enumerate(batches):
                waited = 0.0
                while True:
                    try:
                        got = ready.get(timeout=1.0)
                        break
                    except queue.Empty:
                        waited += 1.0
                        if errors:
                            raise errors[0]
                        if not pt.is_alive():
                            raise RuntimeError(f"the feature reader stopped without delivering batch {k}") from None
                        if reader_timeout_s is not None and waited >= reader_timeout_s:
                            raise RuntimeError(f"the feature reader made no progress for {reader_timeout_s:g} s "
                                               f"(batch {k})") from None
                if got is None or errors:
                    raise errors[0]
                _, spk = got
                a = k & 1
                real = hi - lo
                dummies = (-real) % 8
                lengths = [int(v) for v in rows[lo:hi]]
                off = sum(lengths)
                arenas[a][off:off + dummies].zero_()
                lengths += [1] * dummies
                rows_dev = arenas[a][:off + dummies].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                arena_tokens[a].put(ev)
                noise = torch.randn(real + dummies, h.noise_dim, device=dev, generator=rng)
                plan_frames = -(-max(lengths) // 32) * 32
                packed = eng.forward_ragged(rows_dev, lengths, spk.to(dev), noise, plan_frames, pcm16=True)
                n_valid = off * up
                pcm_buf = pcm_pool.get()
                pcm_buf[:n_valid].copy_(packed[:n_valid], non_blocking=True)
                done = torch.cuda.Event()
                done.record(stream)
                to_write.put((done, pcm_buf, lo, hi))
                stats.utterances += real
                stats.audio_s += n_valid / sampling_rate
                stats.batches += 1
                if errors:
                    raise errors[0]
    except BaseException as e:
        failure = e
    finally:
        # Never leave the helper threads blocked (they hold ~130 MB of pinned arenas): flag the failure,
        # unblock the prefetch thread's put / get, hand the writer its sentinel, join both.
        if failure is not None:
            errors.append(failure)
            for q in arena_tokens:
                q.put(None)
            while pt.is_alive():
                try:
                    ready.get(timeout=0.05)
                except queue.Empty:
                    pass
        to_write.put(None)
        wt.join()
        pt.join()
    if failure is not None:
        raise failure
    if errors:
        raise errors[0]
    torch.cuda.synchronize(dev)
    stats.wall_s = time.perf_counter() - t0
    return stats


def synthesize_files_sharded(gen, items: Sequence[Item], feat_dir: str, spk_emb_dir: str, out_dir: str,
                             rank: int, world: int, readers: int = 4, **kwargs) -> Stats:
    """One process per GPU over a corpus: every rank reads the headers (cheap, one library call), all
    derive the same frame-balanced contiguous partition (`sharding.shard_by_frames`) and each
    synthesises its own range -- independent utterances, no collective.  Per-rank noise streams are
    decorrelated through the seed."""
    from .sharding import shard_by_frames
    rows, _ = feature_shapes([os.path.join(feat_dir, it.name + ".npy") for it in items], readers)
    lo, hi = shard_by_frames(rows, world)[rank]
    kwargs.setdefault("seed", 1234)
    kwargs["seed"] = int(kwargs["seed"]) + rank
    return synthesize_files(gen, items[lo:hi], feat_dir, spk_emb_dir, out_dir, readers=readers, **kwargs)
