"""Benchmark of the Vec2Wav generator forward pass (BASELINE.json metric: audio-seconds generated
per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2]

One "step" = one forward of the workload batch.  At N=1 the workload is BASELINE.json
configs[1] (cfg2: batch 16 x 4 s utterances, 768-ch features, bf16 path, shipped-default
ResBlock2).  N>1 (torchrun, one rank per GPU): every rank generates its own cfg2 batch -- batches
of independent utterances shard with no collective on the hot path ("weak" scaling); NCCL is
used only for the barrier, the max-over-ranks timing and one output gather after the timed
region.

Keys of the JSON line (see the task contract):
  value      whole-job audio-s/s, inputs resident in HBM, per-step CUDA-event times summed over K
             steps, max over ranks; L2 is flushed (256 MB write) before every timed step.
  e2e        same metric through Generator.engine().forward_host -> v2w_forward_host: pinned host
             inputs, H2D + forward + D2H inside the timed region.
  roofline   tensor-bound: algorithmic FLOPs of the tcgen05 tap-GEMM launches / their summed
             CUDA-event durations (events recorded by the library around each launch, extra steps
             after the timed region), against MEASURED_PEAKS.json's sustained bf16 figure.
  cpu_baseline  the oracle's torch-CPU port of the reference forward on this box's host cores,
             bounded sample, rank 0 at N=1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
warnings.filterwarnings("ignore")

CONFIGS = {
    # name: (batch, frames, num_wv_feat)   -- SURVEY.md section 8(d)
    "cfg1": (1, 100, 768),
    "cfg2": (16, 200, 768),
    "cfg3": (32, 300, 1024),
    "cfg4": (1, 1500, 768),
    "cfg5_per_gpu": (64, 500, 768),
}
FRAMES_PER_SECOND = 50
METRIC = "audio-sec/sec Vec2Wav gen"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--resblock", default="2", choices=["1", "2"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def make_hparams(args, feat):
    from wavthruvec_pytorch_b200.hparams import HParams
    h = HParams()
    h.num_wv_feat = feat
    h.resblock = '1' if args.resblock == "1" else 1   # shipped default int 1 => ResBlock2
    return h


def workload_name(args, batch, frames, feat):
    return (f"{args.config}: batch {batch} x {frames / FRAMES_PER_SECOND:g} s utterances "
            f"({frames} frames x {feat}-ch wav2vec features + 192-d spk emb + 192-d noise), "
            f"ResBlock{args.resblock}, random init (seed 1234)")


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's torch port of the reference forward on the host cores
# --------------------------------------------------------------------------------------------
def cpu_forward_timer(args, batch, frames, feat, seconds, steps=None, warmup=2):
    """Times oracle.vec2wav_torch_cpu.generator_forward on a bounded sample of the workload.
    Returns (audio_s_per_s, ms_per_step, sample description, threads, steps done)."""
    import torch
    from oracle import fixtures, vec2wav_torch_cpu
    from wavthruvec_pytorch_b200 import Generator
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use every core this process may run on
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncores))
    h = make_hparams(args, feat)
    gen = fixtures.construct(Generator, h).eval()
    sd = {k: v.detach() for k, v in gen.state_dict().items()}
    sample_b = min(batch, 2)
    x, s, n = fixtures.make_inputs(h, sample_b, frames, 1234)
    threads = torch.get_num_threads()
    for _ in range(warmup):
        vec2wav_torch_cpu.generator_forward(sd, h, x, s, n)
    times = []
    t_end = time.perf_counter() + seconds
    while True:
        t0 = time.perf_counter()
        vec2wav_torch_cpu.generator_forward(sd, h, x, s, n)
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() >= t_end or len(times) >= 200:
            break
    audio = sample_b * frames / FRAMES_PER_SECOND
    mean = sum(times) / len(times)
    sample = (f"{sample_b} of the {batch} utterances of the workload ({frames} frames each), "
              f"{len(times)} forwards, fp32 torch-CPU port of the reference forward "
              f"(oracle/vec2wav_torch_cpu.py), mean {mean * 1e3:.1f} ms, best {min(times) * 1e3:.1f} ms")
    return audio / mean, mean * 1e3, sample, threads, len(times)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def torch_eager_context(gen, x, s, n, audio_per_step, steps=5):
    import torch
    out = {"note": "library context, not this repo's path: Generator._forward_autograd under no_grad "
                   "(same ATen calls as vec2wav/models.py:116-147)"}
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)),
                      ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with torch.no_grad(), ctx:
                for _ in range(2):
                    gen._forward_autograd(x, s, n)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(steps):
                    gen._forward_autograd(x, s, n)
                b.record()
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            out[name] = {"ms_per_step": ms, "value": audio_per_step / ms * 1e3}
        except Exception as e:  # context only: never fail the bench line over it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, frames, feat = CONFIGS[args.config]
    import torch
    v, ms, sample, threads, done = cpu_forward_timer(args, batch, frames, feat, seconds=1e9,
                                                     steps=args.steps, warmup=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, batch, frames, feat),
                   "note": "host cores only; each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpu_count": os.cpu_count()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1417.3), d.get("hbm_gbs", 6453.1), "measured"
    return 1590.0, 6650.0, "fallback"


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
def run_native_arm(args):
    import torch
    import torch.distributed as dist
    from wavthruvec_pytorch_b200 import Generator, native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run "
                             "(one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    batch, frames, feat = CONFIGS[args.config]
    h = make_hparams(args, feat)
    torch.manual_seed(1234)                      # hp.seed (vec2wav/hparams.py:37)
    gen = Generator(h).eval().to(dev)
    gen.precision = args.precision
    eng = gen.engine()
    # distinct utterances per rank (seed offset); same shapes => weak scaling
    rng = torch.Generator().manual_seed(1234 + rank)
    x = torch.randn(batch, feat, frames, generator=rng)
    s = torch.randn(batch, h.spk_dim, generator=rng)
    n = torch.randn(batch, h.noise_dim, generator=rng)
    xd, sd_, nd = x.to(dev), s.to(dev), n.to(dev)
    xh, sh, nh = x.pin_memory(), s.pin_memory(), n.pin_memory()
    out_h = torch.empty((batch, 1, frames * 320), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    audio_per_step = batch * frames / FRAMES_PER_SECOND

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            flush.zero_()
            y = gen(xd, sd_, nd)
        torch.cuda.synchronize(dev)
        launches_per_step = eng.last_launches

        # ---- value: device-resident inputs, per-step events, L2 flushed before each step ----
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()
            starts[i].record()
            y = gen(xd, sd_, nd)
            ends[i].record()
        barrier()
        wall = time.perf_counter() - wall0
        step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
        total_ms = sum(step_ms)

        # ---- e2e: pinned host buffers through the C-ABI.  Every step copies its inputs H2D and its
        # whole waveform D2H inside the timed region.  Two flavours: one blocking v2w_forward_host
        # call per step, and the streaming v2w_pipeline_* API (same copies, overlapped with the
        # kernels of the neighbouring steps) -- the latter is the throughput figure reported as e2e.
        for _ in range(3):
            eng.forward_host(xh, sh, nh, out_h, dev)
        e2e_steps = max(5, min(args.steps, 50))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.forward_host(xh, sh, nh, out_h, dev)
        torch.cuda.synchronize(dev)
        e2e_sync_s = time.perf_counter() - t0
        pipe = eng.host_pipeline(batch, frames, dev)
        out_h2 = torch.empty_like(out_h).pin_memory()
        for i in range(4):
            pipe.submit(xh, sh, nh, out_h if i % 2 == 0 else out_h2)
        pipe.wait()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            pipe.submit(xh, sh, nh, out_h if i % 2 == 0 else out_h2)
        pipe.wait()
        e2e_s = time.perf_counter() - t0
        pipe.close()
        clocks = sampler.stop()

        # ---- roofline: per-op events on the tcgen05 tap-GEMM launches ----
        prof = eng.profile_ops(xd, sd_, nd, reps=5, before_each=lambda: flush.zero_())

    t = torch.tensor([total_ms, e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # output gather (outside the timed region): the only data collective of the job
        g0 = time.perf_counter()
        gathered = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
        dist.gather(y, gathered, dst=0)
        torch.cuda.synchronize(dev)
        gather_ms = (time.perf_counter() - g0) * 1e3
    else:
        gather_ms = 0.0
    total_ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * audio_per_step / (ms_per_step * 1e-3)
        e2e_value = world * audio_per_step / (e2e_s / e2e_steps)
        peak_tf, peak_gbs, peak_src = measured_peaks()
        gemm = [(op, ms) for op, ms in prof if op.kind in (native.OP_TAPGEMM, native.OP_MRF_FUSED, native.OP_PAIR_FUSED)]
        gemm_flops = sum(op.flops for op, _ in gemm)
        gemm_ms = sum(ms for _, ms in gemm)
        all_ms = sum(ms for _, ms in prof)
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            try:
                traffic = json.loads(tpath.read_text()).get(f"{args.config}_rb{args.resblock}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(args, batch, frames, feat),
                       "per_gpu_batch": batch, "l2": "flushed with a 256 MB write before every step",
                       "precision": args.precision, "sharding": f"dp{world} (independent utterances)"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(xh.numel() + sh.numel() + nh.numel()) * 4,
                    "d2h_bytes_per_step": int(out_h.numel()) * 4, "steps": e2e_steps,
                    "api": "Generator.engine().host_pipeline -> v2w_pipeline_submit/wait (2 steps in flight)",
                    "note": "every step copies its inputs H2D and its waveform D2H inside the timed region; no L2 flush "
                            "between these steps (the inputs arrive through the copy), which is why it can edge past `value`",
                    "blocking_call_value": world * audio_per_step / (e2e_sync_s / e2e_steps)},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "tapgemm_simt_kernel" if args.precision == "fp32" else "tapgemm_tc_kernel + mrf_stacked_kernel + mrf_stage_kernel",
                         "launches": len(gemm), "gemm_ms_per_step": gemm_ms,
                         "all_ops_ms_per_step": all_ms,
                         "algorithmic_gflop_per_step": gemm_flops / 1e9},
            "clocks": clocks,
            "wall_s_timed_region": wall,
            "gather_ms": gather_ms,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, ms, sample, threads, _ = cpu_forward_timer(args, batch, frames, feat, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": sample, "host_cpu_count": os.cpu_count(),
                                    "cpu_model": cpu_model()}
            # Context only: the reference's own op sequence -- stock PyTorch/cuDNN ops,
            # weight-norm hooks re-run every call -- on this same GPU, same inputs.  It is the
            # drop-in's train()/autograd delegate, bit-identical to the reference on CPU.
            line["torch_eager_gpu"] = torch_eager_context(gen, xd, sd_, nd, audio_per_step)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
