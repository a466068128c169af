"""Side-stream staging of the next sparse batch: host->device copies plus the whole coordinate manager (cuckoo
coordinate map, strided maps, kernel maps) are built on a second CUDA stream while the current step runs.

Why: `SparseTensor(features, coordinates)` needs one host sync per coordinate map (the row count sizes the next buffers).
Issued on the training stream, that sync waits for the previous step's whole backlog and leaves the GPU idle until the
host has queued new work (~5 ms of a 21 ms step).  On a side stream the sync waits only for the ~1 ms of map kernels and
the training stream never drains.  This is the device-side analogue of the reference's DataLoader workers
(lib/dataset.py:337-416) + the batch transfer Lightning does before `model_step` (pl_BaselineTrainer.py:288-300).

    pf = SparseBatchPrefetcher(device)
    ticket = pf.stage(coords, feats, labels)          # host (pinned) or device tensors
    for ...:
        sinput, target = pf.get(ticket)               # training stream waits on the staging event (no host sync)
        ticket = pf.stage(next_coords, next_feats, next_labels)
        ... forward / backward / step on sinput ...
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

from . import minkowski as E


class SparseBatchPrefetcher:
    """Building a coordinate manager blocks the host five times (each coordinate map's row count sizes the next buffers).
    `high_priority=True` gives the staging stream CUDA's high priority so that its tiny kernels are scheduled ahead of the
    training stream's queued CTAs (staging call 2-4 ms -> 1.2 ms of host time), but the end-to-end step became erratic with
    it (16.1-19.4 ms vs a steady 16.1 ms), hence off by default.  `threaded=True` moves `stage` to a worker thread (the waits happen inside the C
    library with the GIL released); measured neutral on a host-bound step (the worker's Python work contends for the
    GIL: forward issue 6.3 -> 9.3 ms), hence off by default."""

    def __init__(self, device=None, feature_dtype=torch.float32, threaded=False, high_priority=False):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(self.device, priority=-1 if high_priority else 0)
        self.feature_dtype = feature_dtype
        self.pool = ThreadPoolExecutor(1, thread_name_prefix="lgs-stage") if threaded else None
        self._tls = threading.local()

    def _stage(self, coords, feats, labels):
        if not getattr(self._tls, "ready", False):       # a new thread starts on device 0
            torch.cuda.set_device(self.device)
            self._tls.ready = True
        with torch.cuda.stream(self.stream):
            c = coords.to(self.device, non_blocking=True)
            f = feats.to(self.device, non_blocking=True).to(self.feature_dtype)
            lab = labels.to(self.device, non_blocking=True) if labels is not None else None
            st = E.SparseTensor(f, c)                  # hash + every learned strided / kernel map, on this stream
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return st, lab, ev

    def stage(self, coords, feats, labels=None):
        """Enqueue copies + map builds on the staging stream; returns a ticket for `get`.  Blocks the calling thread only
        when `threaded` is off (then for the staging stream's own work)."""
        if self.pool is not None:
            return self.pool.submit(self._stage, coords, feats, labels)
        return self._stage(coords, feats, labels)

    def get(self, ticket):
        """Hand the staged batch to the current (training) stream."""
        st, lab, ev = ticket.result() if self.pool is not None else ticket
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        # the tensors were allocated on the staging stream: tell the caching allocator the training stream uses them
        for t in st.coordinate_manager.tensors():
            t.record_stream(cur)
        for t in (st.F, lab, getattr(st, "unique_index", None), getattr(st, "inverse_mapping", None)):
            if t is not None:
                t.record_stream(cur)
        return st, lab
