"""ScanFitter: the whole hot path (network forward -> tightness vectors / labels -> markers -> LM fit -> SMPL mesh) as one
callable, replayed from CUDA graphs, with several batches in flight.

The ~1300 kernel launches of one step are all static-shape and stream-ordered, so after two eager warm-up calls per input
shape the step is captured (torch.cuda.graphs; every C-ABI entry point launches on the capturing stream) and later calls
only copy the scans into a graph's static input buffer and replay it: no Python, allocator or launch overhead on the
critical path.

A step has long latency-bound stretches that leave most of the GPU idle (furthest point sampling: 1 CTA per scan for
2.5 ms; the LM fit: 1 CTA per scan for 6 ms; the deep PointTransformer levels).  `in_flight` > 1 keeps that many
independent copies of the graph (own buffers, own stream) and `submit()` rotates through them, so batch i+1's network
runs on the SMs batch i's fit leaves idle: +30 % scans/s on a B200 at in_flight=3.  `submit()` returns a Ticket;
`Ticket.result()` makes the caller's stream wait for that batch and returns its output dict.  The outputs are the slot's
static tensors: they are overwritten when the slot is reused, `in_flight` submits later -- `submit()` orders that replay
after whatever the caller has queued on its current stream at that moment, so reads (clones, D2H copies) issued on the
caller's stream before the re-submit are safe; reads issued later, or on other streams, are the caller's to order.
`fitter(pts)` is `submit(pts).result()`.
"""
import torch

from .models import fit_SMPL


class Ticket:
    """One submitted batch: `stream` is the CUDA stream it runs on, `done` the event recorded after its last kernel."""

    def __init__(self, out, stream, done):
        self._out, self.stream, self.done = out, stream, done

    def result(self):
        """Order the caller's current stream after this batch and return dict(vertices [B,6890,3], joints [B,45,3],
        params [B,85], markers, valid, labels, tightness, inner, confidences, iters, errs)."""
        if self.done is not None:
            torch.cuda.current_stream(self.stream.device).wait_event(self.done)
        return self._out


class _Slot:
    __slots__ = ("graph", "static_in", "static_out", "stream", "done")


class ScanFitter:
    def __init__(self, net, args, gender="neutral", use_graph=True, scale_magnitude=10.0,
                 steps_stage0=30, steps_stage1=50, lr_stage0=0.5, lr_stage1=0.2, in_flight=1, sm_budget=None):
        self.net, self.args, self.gender = net, args, gender
        self.use_graph = use_graph
        self.in_flight = max(1, int(in_flight)) if use_graph else 1
        self.scale = scale_magnitude
        self.lm = dict(steps_stage0=steps_stage0, steps_stage1=steps_stage1, lr_stage0=lr_stage0, lr_stage1=lr_stage1)
        self.sm_budget = sm_budget   # SMs the persistent kernels fill; None = SM count - scans per batch when batches overlap
        self._slots = {}     # (shape, device index) -> [ _Slot ] * in_flight
        self._next = {}
        self.launches_per_step = None

    def _step(self, pts):
        out, _ = self.net(pts, ["confidence", "direction", "magnitude"], "standard_vector")
        labels, vec, inner = self.net.postprocess(pts, out, self.scale)
        markers, valid = fit_SMPL.get_markers(self.args, inner, labels, out["confidences"])
        tables = fit_SMPL.body_tables(self.args, self.gender, pts.device)
        fit = fit_SMPL.lm_fit(tables, markers, valid, **self.lm)
        fit.update(markers=markers, valid=valid, labels=labels, tightness=vec, inner=inner, confidences=out["confidences"])
        return fit

    def _build_slots(self, key, example):
        """example: CUDA tensor of the shape to capture for."""
        from . import _lib
        dev = example.device
        warm_in = example.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up: builds plans, caches, lazily prepared tensor-core weights
            for _ in range(2):
                self._step(warm_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.in_flight > 1:
            # leave one SM per scan to the one-CTA-per-scan kernels (FPS, LM fit) of the other batches in flight: the persistent
            # full-grid kernels then never wait for an SM one of those holds for milliseconds (+2 % scans/s at 8 batches in flight)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            _lib.lib().etch_set_sm_budget(int(self.sm_budget) if self.sm_budget else max(sms // 2, sms - int(example.shape[0])))
        slots = []
        for _ in range(self.in_flight):
            sl = _Slot()
            sl.static_in = warm_in.clone()
            sl.stream = torch.cuda.Stream(device=dev)
            sl.done = torch.cuda.Event()
            sl.graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(sl.graph):
                sl.static_out = self._step(sl.static_in)
            self.launches_per_step = _lib.launch_count - l0
            slots.append(sl)
        torch.cuda.synchronize()
        self._slots[key] = slots
        self._next[key] = 0
        return slots

    @torch.no_grad()
    def submit(self, pts, device=None):
        """pts [B,N,3] float32: a CUDA tensor, or a (pinned) host tensor together with `device` -- the host->device copy then
        goes straight into a graph's static input buffer on that batch's stream.  Returns a Ticket.  Everything is launched
        on the device that owns `pts` (or `device`), whatever the thread's current device is."""
        if not pts.is_cuda:
            if device is None:
                raise RuntimeError("etch_b200 has no CPU path: pass a CUDA tensor, or a host tensor plus the target CUDA device")
            dev = torch.device(device)
            if dev.type != "cuda":
                raise RuntimeError("etch_b200 has no CPU path: device must be a CUDA device")
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
        else:
            dev = pts.device
        with torch.cuda.device(dev):
            return self._submit_on(pts, dev)

    def _submit_on(self, pts, dev):
        if not self.use_graph:
            return Ticket(self._step(pts if pts.is_cuda else pts.to(dev, non_blocking=True)), torch.cuda.current_stream(dev), None)
        key = (tuple(pts.shape), dev.index)
        slots = self._slots.get(key)
        if slots is None:
            slots = self._build_slots(key, pts if pts.is_cuda else pts.to(dev))
        i = self._next[key]
        self._next[key] = (i + 1) % len(slots)
        sl = slots[i]
        # Order the slot's stream after everything the caller has queued so far on its current stream of that device: the
        # producer of a CUDA `pts`, and -- for host input too -- any clone / D2H copy of this slot's previous outputs
        # (they are static tensors that the replay below overwrites).
        sl.stream.wait_stream(torch.cuda.current_stream(dev))
        if pts.is_cuda:
            pts.record_stream(sl.stream)
        with torch.cuda.stream(sl.stream):
            sl.static_in.copy_(pts, non_blocking=True)
            sl.graph.replay()
            sl.done.record(sl.stream)
        return Ticket(sl.static_out, sl.stream, sl.done)

    def __call__(self, pts, device=None):
        return self.submit(pts, device).result()
