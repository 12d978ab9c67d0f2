"""Host side of the native generation path: parameter folding cache, plan cache, launch.

PyTorch is plumbing here -- device memory (`torch.empty`), the current CUDA stream and the
module's parameter storage; all arithmetic happens in libv2w (include/v2w.h).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from . import native


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


class NativeEngine:
    """Owns the libv2w model/plan handles of one Generator instance."""

    def __init__(self, module):
        self._module = module
        self._lib = native.load()
        self._cfg = native.make_config(module.h)
        handle = C.c_void_p()
        native.check(self._lib.v2w_model_create(C.byref(self._cfg), C.byref(handle)))
        self._model = handle
        self._packed: Dict[Tuple, torch.Tensor] = {}     # (device, 16-bit format) -> packed weights
        self._packed_sig = None
        # (batch, frames, precision, device, stream) -> (plan handle, workspace), least recently used
        # first.  Bounded by count and by workspace bytes: a validation loop that feeds every utterance
        # at its own length would otherwise keep one ~96 KB-per-frame workspace per distinct length.
        self._plans: "OrderedDict[Tuple, Tuple[C.c_void_p, torch.Tensor]]" = OrderedDict()
        self.max_plans = int(os.environ.get("V2W_MAX_PLANS", "16"))
        self.max_workspace_bytes = int(float(os.environ.get("V2W_MAX_WORKSPACE_GB", "48")) * (1 << 30))
        self._lock = threading.RLock()     # plan cache + folding cache are shared by every caller thread
        self._pinned: Dict[Tuple, int] = {}   # plans a HostPipeline holds on to: never evicted
        self._packers: Dict[torch.device, C.c_void_p] = {}   # device-side fold tables, one per device
        self.last_fold_launches = 0
        self._convs = None        # cached module list / tensor list for the per-call change check
        self._tensors = None
        self._structure = None
        self.last_launches = 0
        self.last_graph_replay = False

    def __del__(self):
        try:
            self._drop_plans()
            for pk in getattr(self, "_packers", {}).values():
                self._lib.v2w_packer_destroy(pk)
            if getattr(self, "_model", None):
                self._lib.v2w_model_destroy(self._model)
                self._model = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _drop_plans(self):
        if getattr(self, "_pinned", None) and any(self._pinned.values()):
            raise RuntimeError("parameters changed while a HostPipeline of this module is open; close() it first")
        for plan, _ws in self._plans.values():
            self._lib.v2w_plan_destroy(plan)
        self._plans.clear()

    def _evict(self, incoming_bytes: int) -> None:
        """Make room for one more plan: least recently used plans go first.  Their workspaces are
        returned to PyTorch's caching allocator, which keeps the memory ordered behind the kernels
        already enqueued on the stream that last used it."""
        def total():
            return sum(ws.numel() for _p, ws in self._plans.values())
        while (len(self._plans) >= self.max_plans or total() + incoming_bytes > self.max_workspace_bytes):
            victim = next((k for k in self._plans if not self._pinned.get(k)), None)
            if victim is None:
                break
            plan, _ws = self._plans.pop(victim)
            self._lib.v2w_plan_destroy(plan)

    def workspace_bytes(self) -> int:
        return sum(ws.numel() for _p, ws in self._plans.values())

    def invalidate(self):
        self._packed.clear()
        self._packed_sig = None
        self._tensors = None
        self._drop_plans()

    def flops_per_frame(self) -> int:
        return int(self._lib.v2w_model_flops_per_frame(self._model))

    def layers(self) -> List[native.LayerDesc]:
        out = []
        for i in range(self._lib.v2w_model_num_layers(self._model)):
            d = native.LayerDesc()
            native.check(self._lib.v2w_model_get_layer(self._model, i, C.byref(d)))
            out.append(d)
        return out

    # ------------------------------------------------------------------------------------------
    # state_dict -> v2w_params -> packed buffer
    # ------------------------------------------------------------------------------------------
    def _state_tensors(self):
        m = self._module
        ts = []
        for conv in m.conv_modules():
            if hasattr(conv, "weight_v"):
                ts += [conv.weight_v, conv.weight_g, conv.bias]
            else:  # after remove_weight_norm()
                ts += [conv.weight, None, conv.bias]
        for fc, cbn in zip(m.fcs, m.cbns):
            ts += [fc.weight, fc.bias, cbn.layer.weight_orig, cbn.layer.weight_u,
                   cbn.layer.weight_v, cbn.layer.bias, cbn.batch_nrom.running_mean,
                   cbn.batch_nrom.running_var]
        return ts

    def _signature(self, ts):
        return tuple((t.data_ptr(), t._version) if t is not None else None for t in ts)

    def pack_host(self, half_format: int = native.HALF_BF16) -> torch.Tensor:
        """Fold + pack on the host (v2w_model_pack_half); returns a uint8 CPU tensor."""
        ts = self._state_tensors()
        host = [None if t is None else t.detach().to("cpu", torch.float32).contiguous() for t in ts]
        nconv = self._lib.v2w_model_num_convs(self._model)
        nstage = self._cfg.num_upsamples
        if len(host) != 3 * nconv + 8 * nstage:
            raise RuntimeError("module structure does not match the native layer table")
        convs = (native.ConvParams * nconv)()
        for i in range(nconv):
            v, g, b = host[3 * i: 3 * i + 3]
            convs[i].v = _ptr(v)
            convs[i].g = _ptr(g) if g is not None else None
            convs[i].bias = _ptr(b)
        cond = (native.CondParams * nstage)()
        base = 3 * nconv
        for i in range(nstage):
            vals = host[base + 8 * i: base + 8 * i + 8]
            for (name, _), t in zip(native.CondParams._fields_, vals):
                setattr(cond[i], name, _ptr(t))
        params = native.Params(nconv, convs, nstage, cond)
        nbytes = int(self._lib.v2w_model_packed_bytes(self._model))
        packed = torch.empty(nbytes, dtype=torch.uint8)
        native.check(self._lib.v2w_model_pack_half(self._model, C.byref(params), packed.data_ptr(),
                                                   half_format))
        del host
        return packed

    def _refresh(self, train: bool = False) -> None:
        """Drop the folded weights (and the plans bound to them) when any parameter or buffer
        changed: storage pointer or in-place version counter (optimizer step, load_state_dict,
        manual edits), or the weight-norm structure (remove_weight_norm on any sub-block).
        Module._apply (.to()/.cuda()/.float()) replaces buffer objects and calls invalidate().
        The tensor list is cached because walking the module tree costs ~4x the check itself.
        `train`: the caller is the train()-mode forward, which reads the convolutions from the packed
        buffer but none of its conditioning section -- only a change of a convolution invalidates."""
        if self._convs is None:
            self._convs = self._module.conv_modules()
        structure = tuple("weight_v" in c._parameters for c in self._convs)
        if self._tensors is None or structure != self._structure:
            self._tensors = self._state_tensors()
            self._structure = structure
        sig = self._signature(self._tensors)
        n = 3 * len(self._convs)
        conv_sig, cond_sig = sig[:n], sig[n:]
        old = self._packed_sig
        if old is None or old[0] != conv_sig or not self._packed:
            if old is not None and self._packed and self._refold_on_device():
                # every packed buffer was re-folded IN PLACE on its device (v2w_packer_run, stream-ordered):
                # plans, workspaces and TMA descriptors stay
                self._packed_sig = (conv_sig, cond_sig)
                return
            self._packed.clear()
            self._drop_plans()   # plans hold pointers into the packed buffer
            self._packed_sig = (conv_sig, cond_sig)
        elif not train and old[1] != cond_sig:
            # only u / v / running statistics / fcs moved (train()-mode forwards, or a caller editing the
            # buffers): re-fold the conditioning section in place, plans and convolutions stay
            self._repack_cond()
            self._packed_sig = (conv_sig, cond_sig)

    # -- device-side fold (v2w_packer_*): no parameter ever leaves the device ----------------------
    def _device_foldable(self, device: torch.device) -> bool:
        if os.environ.get("V2W_HOST_FOLD"):
            return False
        if self._tensors is None:
            self._tensors = self._state_tensors()
        return all(t is None or (t.is_cuda and t.device == device and t.dtype == torch.float32 and t.is_contiguous())
                   for t in self._tensors)

    def _packer(self, device: torch.device):
        pk = self._packers.get(device)
        if pk is None:
            handle = C.c_void_p()
            with torch.cuda.device(device):
                native.check(self._lib.v2w_packer_create(self._model, C.byref(handle)))
            pk = self._packers[device] = handle
        return pk

    def _device_params(self):
        """v2w_params over the module's own device storage (kept alive by the returned tuple)."""
        ts = self._tensors
        nconv = self._lib.v2w_model_num_convs(self._model)
        nstage = self._cfg.num_upsamples
        if len(ts) != 3 * nconv + 8 * nstage:
            raise RuntimeError("module structure does not match the native layer table")
        convs = (native.ConvParams * nconv)()
        for i in range(nconv):
            v, g, b = ts[3 * i: 3 * i + 3]
            convs[i].v = _ptr(v)
            convs[i].g = _ptr(g) if g is not None else None
            convs[i].bias = _ptr(b)
        cond = (native.CondParams * nstage)()
        base = 3 * nconv
        for i in range(nstage):
            for (name, _), t in zip(native.CondParams._fields_, ts[base + 8 * i: base + 8 * i + 8]):
                setattr(cond[i], name, _ptr(t))
        return native.Params(nconv, convs, nstage, cond), convs, cond

    def pack_device(self, device: torch.device, half_format: int = native.HALF_BF16,
                    out: torch.Tensor = None) -> torch.Tensor:
        """Fold + pack on the device (v2w_packer_run) on the current stream of `device`: four launches,
        bit-identical to pack_host.  `out`: an existing packed buffer to fold into (in place)."""
        params, _convs, _cond = self._device_params()
        nbytes = int(self._lib.v2w_model_packed_bytes(self._model))
        if out is None:
            out = torch.empty(nbytes, dtype=torch.uint8, device=device)
        n = C.c_int32(0)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            native.check(self._lib.v2w_packer_run(self._packer(device), C.byref(params), out.data_ptr(),
                                                  half_format, C.c_void_p(stream), C.byref(n)))
        self.last_fold_launches = n.value
        return out

    def _refold_on_device(self) -> bool:
        if not all(self._device_foldable(dev) for (dev, _fmt) in self._packed):
            return False
        for (dev, fmt), buf in self._packed.items():
            self.pack_device(dev, fmt, out=buf)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                for key, (plan, _ws) in self._plans.items():
                    if key[3] == dev and int(self._lib.v2w_plan_packed_format(plan)) == fmt:
                        native.check(self._lib.v2w_plan_refresh_weights(plan, C.c_void_p(stream)))
        return True

    def _repack_cond(self) -> None:
        nstage = self._cfg.num_upsamples
        if all(self._device_foldable(dev) for (dev, _fmt) in self._packed):
            _params, _convs, cond = self._device_params()
            for (dev, _fmt), buf in self._packed.items():
                with torch.cuda.device(dev):
                    stream = torch.cuda.current_stream(dev).cuda_stream
                    native.check(self._lib.v2w_packer_run_cond(self._packer(dev), cond, nstage, buf.data_ptr(),
                                                               C.c_void_p(stream)))
            return
        ts = self._tensors[3 * len(self._convs):]
        host = [t.detach().to("cpu", torch.float32).contiguous() for t in ts]
        cond = (native.CondParams * nstage)()
        for i in range(nstage):
            for (name, _), t in zip(native.CondParams._fields_, host[8 * i: 8 * i + 8]):
                setattr(cond[i], name, _ptr(t))
        off, nbytes = C.c_int64(0), C.c_int64(0)
        native.check(self._lib.v2w_model_cond_region(self._model, C.byref(off), C.byref(nbytes)))
        region = torch.empty(nbytes.value, dtype=torch.uint8)
        native.check(self._lib.v2w_model_pack_cond(self._model, cond, nstage, region.data_ptr()))
        for (device, _fmt), buf in self._packed.items():
            with torch.cuda.device(device):
                # stream-ordered behind every forward already enqueued on the current stream
                buf[off.value: off.value + nbytes.value].copy_(region)

    def _packed_on(self, device: torch.device, half_format: int) -> torch.Tensor:
        buf = self._packed.get((device, half_format))
        if buf is None:
            if self._device_foldable(device):
                buf = self.pack_device(device, half_format)
            else:
                buf = self.pack_host(half_format).to(device)
            self._packed[(device, half_format)] = buf
        return buf

    # ------------------------------------------------------------------------------------------
    def plan(self, batch: int, frames: int, precision: str, device: torch.device, train: bool = False):
        """(plan handle, workspace) for this problem size ON THE CURRENT STREAM of `device`; created and
        bound on first use.  A plan owns a workspace, a side stream and fork / join events, so it serves
        one stream at a time: callers on different streams (or threads using different streams) get
        separate plans; callers sharing a stream are serialised by that stream."""
        with self._lock:
            return self._plan_locked(batch, frames, precision, device, train)

    def _plan_locked(self, batch, frames, precision, device, train):
        # `module.parameters_frozen`: the caller guarantees that no parameter or buffer changes between
        # calls (serving): skip the per-call walk over ~150 (data_ptr, _version) pairs
        if not (getattr(self._module, "parameters_frozen", False) and self._packed_sig is not None and self._packed):
            self._refresh(train)
        with torch.cuda.device(device):
            stream_id = torch.cuda.current_stream(device).cuda_stream
        key = (batch, frames, precision, device, stream_id)
        hit = self._plans.get(key)
        if hit is not None:
            self._plans.move_to_end(key)
            return hit
        prec = native.PRECISIONS[precision]
        handle = C.c_void_p()
        native.check(self._lib.v2w_plan_create(self._model, batch, frames, prec, C.byref(handle)))
        try:
            packed = self._packed_on(device, int(self._lib.v2w_plan_packed_format(handle)))
            ws_bytes = int(self._lib.v2w_plan_workspace_bytes(handle))
            self._evict(ws_bytes)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                native.check(self._lib.v2w_plan_bind(handle, packed.data_ptr(), ws.data_ptr()))
        except Exception:
            self._lib.v2w_plan_destroy(handle)
            raise
        self._plans[key] = (handle, ws)
        return handle, ws

    def ops(self, batch: int, frames: int, precision: str, device) -> List[native.OpDesc]:
        handle, _ = self.plan(batch, frames, precision, torch.device(device))
        out = []
        for i in range(self._lib.v2w_plan_num_ops(handle)):
            d = native.OpDesc()
            native.check(self._lib.v2w_plan_get_op(handle, i, C.byref(d)))
            out.append(d)
        return out

    def launches(self, batch: int, frames: int, precision: str, device) -> List[native.LaunchDesc]:
        """Launch geometry of every op of the (bound) plan: v2w_plan_get_launch."""
        handle, _ = self.plan(batch, frames, precision, torch.device(device))
        out = []
        for i in range(self._lib.v2w_plan_num_ops(handle)):
            d = native.LaunchDesc()
            native.check(self._lib.v2w_plan_get_launch(handle, i, C.byref(d)))
            out.append(d)
        return out

    # ------------------------------------------------------------------------------------------
    def _check_inputs(self, x, spk, noise):
        h = self._module.h
        if x.dim() != 3 or x.shape[1] != h.num_wv_feat:
            raise RuntimeError(f"expected x of shape [B, {h.num_wv_feat}, T], got {tuple(x.shape)}")
        B = x.shape[0]
        if spk.dim() != 2 or tuple(spk.shape) != (B, h.spk_dim):
            raise RuntimeError(f"expected spk_emb of shape [{B}, {h.spk_dim}], got {tuple(spk.shape)}")
        if noise.dim() != 2 or tuple(noise.shape) != (B, h.noise_dim):
            raise RuntimeError(f"expected noise of shape [{B}, {h.noise_dim}], got {tuple(noise.shape)}")

    def forward(self, x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """Device tensors in, device tensor out, asynchronous on the current stream."""
        self._check_inputs(x, spk, noise)
        if not x.is_cuda:
            raise RuntimeError(
                "the native Vec2Wav generation path runs on sm_100 GPUs only and has no CPU "
                "fallback; move the module and inputs to a CUDA device (train() mode / autograd "
                "use the PyTorch delegate on any device)")
        if spk.device != x.device or noise.device != x.device:
            raise RuntimeError("x, spk_emb and noise must live on the same device")
        dev = x.device
        x = x.detach().to(torch.float32).contiguous()
        spk = spk.detach().to(torch.float32).contiguous()
        noise = noise.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        out = torch.empty((B, 1, T * self.total_upsample()), dtype=torch.float32, device=dev)
        # Micro-batches: utterances are independent, so a large batch is run as consecutive launches of at most
        # `max_batch_frames` frames through ONE plan (workspace ~96 KB per frame: 256 x 500 frames would take
        # 12 GB and, measured at the cfg5 shard sizes, run 5 % slower per utterance than 4 x 64 x 500 --
        # the working set of a stage falls out of L2).  Each launch writes its slice of `out` in place.
        limit = getattr(self._module, "max_batch_frames", None)
        if limit and B > 1 and B * T > limit:
            chunks = -(-(B * T) // int(limit))          # as few launches as the limit allows, ...
            mb = max(1, -(-B // chunks))                # ... of equal size (one plan shape, except a ragged tail)
            total = 0
            for lo in range(0, B, mb):
                hi = min(B, lo + mb)
                total += self._launch(x[lo:hi], spk[lo:hi], noise[lo:hi], out[lo:hi], dev)
            self.last_launches = total
            return out
        self.last_launches = self._launch(x, spk, noise, out, dev)
        return out

    def _launch(self, x, spk, noise, out, dev) -> int:
        """One v2w_forward (or graph replay) of contiguous fp32 device tensors into `out`."""
        B, _, T = x.shape
        handle, _ws = self.plan(B, T, self._module.precision, dev)
        n = C.c_int32(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if getattr(self._module, "use_cuda_graph", False):
                # latency path: replay from a CUDA graph keyed by the four device pointers (a serving
                # loop with preallocated buffers, or PyTorch's allocator handing the same blocks back)
                hit = C.c_int32(0)
                native.check(self._lib.v2w_forward_graph(handle, x.data_ptr(), spk.data_ptr(),
                                                         noise.data_ptr(), out.data_ptr(),
                                                         C.c_void_p(stream), C.byref(n), C.byref(hit)))
                self.last_graph_replay = bool(hit.value)
            else:
                native.check(self._lib.v2w_forward(handle, x.data_ptr(), spk.data_ptr(),
                                                   noise.data_ptr(), out.data_ptr(),
                                                   C.c_void_p(stream), C.byref(n)))
        return n.value

    def forward_into(self, x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """forward() into a caller-owned output tensor: with `use_cuda_graph` and preallocated
        fp32 contiguous buffers every call after the second is a single graph launch."""
        self._check_inputs(x, spk, noise)
        dev = x.device
        B, _, T = x.shape
        for t in (x, spk, noise, out):
            if not t.is_cuda or t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("forward_into takes contiguous fp32 CUDA tensors on one device")
        if tuple(out.shape) != (B, 1, T * self.total_upsample()):
            raise RuntimeError("out has the wrong shape")
        handle, _ws = self.plan(B, T, self._module.precision, dev)
        n, hit = C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            fn = self._lib.v2w_forward_graph if getattr(self._module, "use_cuda_graph", False) else None
            if fn is not None:
                native.check(fn(handle, x.data_ptr(), spk.data_ptr(), noise.data_ptr(), out.data_ptr(),
                                C.c_void_p(stream), C.byref(n), C.byref(hit)))
            else:
                native.check(self._lib.v2w_forward(handle, x.data_ptr(), spk.data_ptr(), noise.data_ptr(),
                                                   out.data_ptr(), C.c_void_p(stream), C.byref(n)))
        self.last_graph_replay = bool(hit.value)
        self.last_launches = n.value
        return out

    def forward_train(self, x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """train()-mode forward under no_grad through v2w_forward_train:
        batch-statistic BN and one spectral-norm power iteration per stage on the device, with the
        reference's side effects -- `running_mean` / `running_var`, `weight_u` / `weight_v` updated in
        place, `num_batches_tracked` incremented (vec2wav/modules.py:14,16 with module.training set)."""
        self._check_inputs(x, spk, noise)
        if not x.is_cuda or spk.device != x.device or noise.device != x.device:
            raise RuntimeError("forward_train takes CUDA tensors on one device (no CPU fallback)")
        m = self._module
        dev = x.device
        x = x.detach().to(torch.float32).contiguous()
        spk = spk.detach().to(torch.float32).contiguous()
        noise = noise.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        st = native.TrainState()
        st.num_stages = len(m.cbns)
        momentum = m.cbns[0].batch_nrom.momentum
        if momentum is None:
            raise RuntimeError("BatchNorm1d(momentum=None) (cumulative average) is not supported natively")
        st.momentum = float(momentum)
        counters = []
        for i, (fc, cbn) in enumerate(zip(m.fcs, m.cbns)):
            bn, lin = cbn.batch_nrom, cbn.layer
            if bn.momentum != momentum:
                raise RuntimeError("all stages must share one BatchNorm momentum")
            ts = (fc.weight, fc.bias, lin.weight_orig, lin.bias, lin.weight_u, lin.weight_v,
                  bn.running_mean, bn.running_var)
            for (name, _), t in zip(native.TrainStage._fields_, ts):
                if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError(f"stage {i} {name}: expected a contiguous fp32 tensor on {dev}")
                setattr(st.stage[i], name, t.data_ptr())
            counters.append(bn.num_batches_tracked)
        handle, _ws = self.plan(B, T, m.train_precision, dev, train=True)
        out = torch.empty((B, 1, T * self.total_upsample()), dtype=torch.float32, device=dev)
        n = C.c_int32(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            native.check(self._lib.v2w_forward_train(handle, x.data_ptr(), spk.data_ptr(),
                                                     noise.data_ptr(), out.data_ptr(), C.byref(st),
                                                     C.c_void_p(stream), C.byref(n)))
        with torch.no_grad():
            torch._foreach_add_(counters, 1)
        # u, v and the running statistics changed behind autograd's version counters: the folded
        # conditioning section is stale for eval(); the convolutions are not
        self._packed_sig = (self._packed_sig[0], None)
        self.last_launches = n.value
        return out

    def halo_frames(self) -> int:
        """Input frames past an utterance's end that can reach its last valid output sample."""
        return int(self._lib.v2w_model_halo_frames(self._model))

    def forward_ragged(self, x_rows: torch.Tensor, lengths, spk: torch.Tensor, noise: torch.Tensor,
                       plan_frames: int, pcm16: bool = False) -> torch.Tensor:
        """Variable-length batch through v2w_forward_ragged.  `x_rows` = the utterances' [T_b, C]
        feature rows concatenated (device, fp32), `lengths` = T_b per utterance (host ints).
        Returns ONE packed waveform buffer (fp32, or int16 PCM): utterance b owns samples
        [off_b * up, (off_b + T_b) * up) with off = exclusive prefix sum of `lengths`."""
        lengths = [int(v) for v in lengths]
        B = len(lengths)
        if not x_rows.is_cuda:
            raise RuntimeError("the native Vec2Wav generation path runs on sm_100 GPUs only and has "
                               "no CPU fallback")
        if x_rows.dim() != 2 or x_rows.shape[1] != self._module.h.num_wv_feat or x_rows.shape[0] != sum(lengths):
            raise RuntimeError(f"x_rows must be [sum(lengths), {self._module.h.num_wv_feat}]")
        if tuple(spk.shape) != (B, self._module.h.spk_dim) or tuple(noise.shape) != (B, self._module.h.noise_dim):
            raise RuntimeError("spk_emb / noise must be [len(lengths), dim]")
        dev = x_rows.device
        x_rows = x_rows.detach().to(torch.float32).contiguous()
        spk = spk.detach().to(device=dev, dtype=torch.float32).contiguous()
        noise = noise.detach().to(device=dev, dtype=torch.float32).contiguous()
        handle, _ws = self.plan(B, int(plan_frames), self._module.precision, dev)
        offs = (C.c_int32 * (B + 1))()
        for b, n_b in enumerate(lengths):
            offs[b + 1] = offs[b] + n_b
        out = torch.empty(offs[B] * self.total_upsample(),
                          dtype=torch.int16 if pcm16 else torch.float32, device=dev)
        n = C.c_int32(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            native.check(self._lib.v2w_forward_ragged(
                handle, x_rows.data_ptr(), offs, spk.data_ptr(), noise.data_ptr(), out.data_ptr(),
                native.OUT_PCM16 if pcm16 else native.OUT_F32, C.c_void_p(stream), C.byref(n)))
        self.last_launches = n.value
        return out

    def forward_host(self, x: torch.Tensor, spk: torch.Tensor, noise: torch.Tensor,
                     out: torch.Tensor, device) -> torch.Tensor:
        """Host (ideally pinned) tensors in and out through v2w_forward_host: H2D, forward, D2H,
        stream synchronised on return."""
        self._check_inputs(x, spk, noise)
        for t in (x, spk, noise, out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("forward_host takes contiguous fp32 CPU tensors")
        dev = torch.device(device)
        B, _, T = x.shape
        if tuple(out.shape) != (B, 1, T * self.total_upsample()):
            raise RuntimeError("out has the wrong shape")
        handle, _ws = self.plan(B, T, self._module.precision, dev)
        n = C.c_int32(0)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            native.check(self._lib.v2w_forward_host(handle, x.data_ptr(), spk.data_ptr(),
                                                    noise.data_ptr(), out.data_ptr(),
                                                    C.c_void_p(stream), C.byref(n)))
        self.last_launches = n.value
        return out

    def host_pipeline(self, batch: int, frames: int, device) -> "HostPipeline":
        """Streaming generation from pinned host buffers (v2w_pipeline_*): copies of neighbouring
        steps overlap the kernels of the current one."""
        dev = torch.device(device)
        with self._lock:
            handle, _ws = self.plan(batch, frames, self._module.precision, dev)
            key = next(k for k, v in self._plans.items() if v[0] is handle)
            self._pinned[key] = self._pinned.get(key, 0) + 1
        return HostPipeline(self, handle, batch, frames, dev, key)

    def profile_ops(self, x, spk, noise, reps: int = 3, before_each=None):
        """Per-op device times (ms, mean over `reps` forwards) from CUDA events recorded by the
        library on the launch stream.  Returns [(OpDesc, ms)]."""
        B, _, T = x.shape
        dev = x.device
        handle, _ws = self.plan(B, T, self._module.precision, dev)
        ops = self.ops(B, T, self._module.precision, dev)
        native.check(self._lib.v2w_plan_set_profiling(handle, 1))
        acc = [0.0] * len(ops)
        try:
            for _ in range(reps):
                if before_each is not None:
                    before_each()
                out = torch.empty((B, 1, T * self.total_upsample()), dtype=torch.float32, device=dev)
                self._launch(x.detach().to(torch.float32).contiguous(), spk.detach().to(torch.float32).contiguous(),
                             noise.detach().to(torch.float32).contiguous(), out, dev)   # ONE launch of this plan (no micro-batches)
                torch.cuda.synchronize(dev)
                for i in range(len(ops)):
                    ms = float(self._lib.v2w_plan_op_time_ms(handle, i))
                    if ms < 0:
                        raise native.V2WError(-3, native.last_error())
                    acc[i] += ms
        finally:
            native.check(self._lib.v2w_plan_set_profiling(handle, 0))
        return [(op, a / reps) for op, a in zip(ops, acc)]

    def total_upsample(self) -> int:
        r = 1
        for u in self._module.h.upsample_rates:
            r *= int(u)
        return r


class HostPipeline:
    """Handle on a v2w_pipeline.  `submit` is asynchronous; `wait` drains and returns the number of
    kernels launched.  Host tensors must be contiguous fp32 CPU tensors (pinned for overlap) and
    stay alive until `wait` returns."""

    def __init__(self, engine: NativeEngine, plan_handle, batch, frames, device, plan_key=None):
        self._engine = engine
        self._plan_key = plan_key
        self._lib = engine._lib
        self._device = device
        self._shape = (batch, frames)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            native.check(self._lib.v2w_pipeline_create(plan_handle, C.byref(handle)))
        self._handle = handle
        self._keep = []

    def submit(self, x, spk, noise, out) -> None:
        self._engine._check_inputs(x, spk, noise)
        for t in (x, spk, noise, out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("HostPipeline takes contiguous fp32 CPU tensors")
        B, T = self._shape
        if x.shape[0] != B or x.shape[2] != T or tuple(out.shape) != (B, 1, T * self._engine.total_upsample()):
            raise RuntimeError("tensor shapes do not match the pipeline's plan")
        self._keep.append((x, spk, noise, out))
        with torch.cuda.device(self._device):
            native.check(self._lib.v2w_pipeline_submit(self._handle, x.data_ptr(), spk.data_ptr(),
                                                       noise.data_ptr(), out.data_ptr()))

    def wait(self) -> int:
        n = C.c_int64(0)
        with torch.cuda.device(self._device):
            native.check(self._lib.v2w_pipeline_wait(self._handle, C.byref(n)))
        self._keep.clear()
        return int(n.value)

    def close(self) -> None:
        if self._handle:
            self._lib.v2w_pipeline_destroy(self._handle)
            self._handle = None
            if self._plan_key is not None:
                with self._engine._lock:
                    self._engine._pinned[self._plan_key] = max(0, self._engine._pinned.get(self._plan_key, 1) - 1)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
