"""Variable-length batched generation + un-padding.

The reference batches utterances by zero-padding their wav2vec features to the longest one
(`collate_fn_tensor`, vec2wav/dataset.py:194-228; `pad_2D_tensor`, vec2wav/utils.py:60-77) and the
generator then synthesises the padded tail of every shorter utterance as well.  Here a batch is

    feats[b] : [T_b, num_wv_feat]   the layout a dataset item already has (dataset.py:181-187)

and generation runs the whole batch through `v2w_forward_ragged` at plan frames = longest utterance:

  * the feature rows are gathered straight from one packed row buffer, with zeros past each
    utterance's end -- exactly the reference's zero padding, without the [B, C, T] detour;
  * every row-tiled kernel skips the tiles that lie past  T_b + halo  frames of their utterance
    (`halo` = the generator's right receptive field, `v2w_model_halo_frames`, 10 frames for the
    shipped configuration): rows out there cannot reach a valid sample;
  * conv_post writes only the valid samples, packed, as fp32 or 16-bit PCM.

Optionally the utterances are first grouped into LENGTH BUCKETS (`max_buckets` > 1; exact dynamic
programme over the sorted lengths) and each bucket runs at its own, shorter plan length.  With tile
skipping in the kernels this is only worth it to bound the workspace (B x longest) of very skewed
batches: measured on B200, 64 utterances of 1-10 s, one skipping launch sequence beats both the
zero-padded batch and four bucketed launches (profiles/r01_varlen.json).

Every returned sample equals what the reference computes for the zero-padded batch and then
truncates (pinned in `tests/test_varlen.py`).  Host logic only; the
arithmetic is libv2w's.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

# One more launch sequence costs about as much as this many frames of one utterance (19 kernels of
# a few microseconds each against ~1.2 us per frame-utterance at cfg2 rates).  Only the bucket
# count depends on it.
LAUNCH_COST_FRAMES = 64
MAX_UTTERANCES_PER_CALL = 512   # simt::kMaxRaggedBatch: row offsets travel as a kernel parameter


def plan_buckets(lengths: Sequence[int], halo: int, max_buckets: int = 4,
                 launch_cost: int = LAUNCH_COST_FRAMES, round_to: int = 8) -> List[Tuple[List[int], int]]:
    """Partition utterances into <= max_buckets groups of similar length.

    Returns [(indices, plan_frames)], longest bucket first; `indices` refer to `lengths`.
    plan_frames = min(longest of the bucket + halo rounded up to `round_to` frames (fewer distinct
    plans to build), longest of the whole batch).  Minimises  sum_b len(bucket) * plan_frames + launch_cost * buckets  exactly (O(n^2 k))."""
    n = len(lengths)
    if n == 0:
        return []
    if min(lengths) < 1:
        raise ValueError("every utterance needs at least one frame")
    order = sorted(range(n), key=lambda i: -int(lengths[i]))
    ls = [int(lengths[i]) for i in order]
    t_max = ls[0]

    def frames(longest: int) -> int:
        # The reference's padded region ends at the batch's longest utterance: a plan may be any
        # length in [longest + halo, t_max], or exactly t_max -- never longer (the longest
        # utterance must see the END of the sequence there, not more zero-feature frames).
        t = longest + halo
        if t >= t_max:
            return t_max
        return min(-(-t // round_to) * round_to, t_max) if round_to > 1 else t

    k_max = max(1, min(max_buckets, n))
    INF = float("inf")
    # best[k][i]: cheapest cover of the first i (longest) utterances with k buckets
    best = [[INF] * (n + 1) for _ in range(k_max + 1)]
    cut = [[0] * (n + 1) for _ in range(k_max + 1)]
    best[0][0] = 0.0
    for k in range(1, k_max + 1):
        for i in range(1, n + 1):
            for j in range(k - 1, i):
                if best[k - 1][j] == INF:
                    continue
                c = best[k - 1][j] + (i - j) * frames(ls[j]) + launch_cost
                if c < best[k][i]:
                    best[k][i], cut[k][i] = c, j
    k = min(range(1, k_max + 1), key=lambda kk: best[kk][n])
    out, i = [], n
    while k > 0:
        j = cut[k][i]
        out.append((order[j:i], frames(ls[j])))
        i, k = j, k - 1
    return out[::-1]


def padded_frames(lengths: Sequence[int], buckets) -> Tuple[int, int]:
    """(frames the reference computes for this batch, frames the bucketed plan computes)."""
    return len(lengths) * max(lengths), sum(len(idx) * t for idx, t in buckets)


@torch.no_grad()
def generate_varlen(gen, feats: Sequence[torch.Tensor], spk_emb: torch.Tensor, noise: torch.Tensor,
                    pcm16: bool = False, max_buckets: int = 1) -> List[torch.Tensor]:
    """Generate one waveform per utterance.

    feats[b]: [T_b, num_wv_feat] float tensor on the generator's CUDA device; spk_emb / noise:
    [B, dim].  Returns B tensors of T_b * 320 samples (fp32 in (-1, 1), or int16 PCM), each a view
    into its bucket's packed output buffer."""
    if gen.training:
        raise RuntimeError("generate_varlen is the eval()-mode generation path")
    eng = gen.engine()
    lengths = [int(f.shape[0]) for f in feats]
    if len(lengths) != spk_emb.shape[0] or len(lengths) != noise.shape[0]:
        raise RuntimeError("feats, spk_emb and noise disagree on the batch size")
    up = eng.total_upsample()
    result: List[torch.Tensor] = [None] * len(lengths)
    if len(lengths) > MAX_UTTERANCES_PER_CALL:
        # v2w_forward_ragged takes at most 512 utterances per call: consecutive chunks, each padded
        # (and tile-skipped) like a batch of its own
        for a in range(0, len(lengths), MAX_UTTERANCES_PER_CALL):
            b = min(a + MAX_UTTERANCES_PER_CALL, len(lengths))
            result[a:b] = generate_varlen(gen, feats[a:b], spk_emb[a:b], noise[a:b], pcm16, max_buckets)
        return result
    for idx, plan_t in plan_buckets(lengths, eng.halo_frames(), max_buckets):
        rows = torch.cat([feats[i] for i in idx], dim=0)
        sel = torch.as_tensor(idx, device=spk_emb.device)
        packed = eng.forward_ragged(rows, [lengths[i] for i in idx], spk_emb.index_select(0, sel),
                                    noise.index_select(0, sel), plan_t, pcm16=pcm16)
        off = 0
        for i in idx:
            result[i] = packed[off * up:(off + lengths[i]) * up]
            off += lengths[i]
    return result


def write_wav(path, pcm: torch.Tensor, sampling_rate: int = 16000) -> None:
    """16-bit mono RIFF/WAVE file from an int16 waveform (host copy of one generate_varlen result)."""
    import struct
    data = pcm.detach().to("cpu", torch.int16).contiguous().numpy().tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt ")
        f.write(struct.pack("<IHHIIHH", 16, 1, 1, sampling_rate, sampling_rate * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
