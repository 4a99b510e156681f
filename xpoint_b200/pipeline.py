"""Sharded pair pipeline: the product form of the batched evaluation driver.

The reference drives inference one DataLoader batch at a time on one GPU, with per-sample Python loops and device<->host
copies in between (xpoint/utils/benchmark_evaluation.py:832-931); its only multi-GPU mechanism is nn.DataParallel for
training.  Pairs are independent end to end, so here a host list of pairs is batch-sharded over the GPUs with no collective
on the data path:

  PairStream            one GPU: pinned double-buffered uploads on a copy stream, the step replayed from a CUDA graph, results
                        downloaded behind the step's kernels; the host blocks one step late (bench.py's e2e number is this loop)
  ShardedPairPipeline   all visible GPUs of one process (a worker thread + PairStream + weight replica per device), or -- when
                        torch.distributed is initialised -- this rank's shard with a host-side gather of the results
"""
from __future__ import annotations

import copy
import threading
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from .sharding import shard_range
from .xpoint import PairPipeline

RESULT_FIELDS = ("kp_optical", "kp_thermal", "n_optical", "n_thermal", "match_idx", "match_dist", "n_matches")


class PairStream:
    """Streams host batches through one PairPipeline on one device.  ``run(batches)`` takes an iterable of
    (optical, thermal) HOST tensors of one fixed shape (B, 1, H, W) (pinned memory makes the uploads asynchronous) and
    returns one dict of host tensors per batch (RESULT_FIELDS [+ H / inliers / n_inliers])."""

    def __init__(self, pipe: PairPipeline, device, use_graph: bool = True, fields: Sequence[str] = RESULT_FIELDS, graphed=None):
        """graphed: an existing GraphedPairPipeline of `pipe` for the batch shape (otherwise captured on first use)."""
        self.pipe, self.device, self.use_graph = pipe, torch.device(device), use_graph or graphed is not None
        self.fields = tuple(fields) + (("H", "inliers", "n_inliers") if pipe.estimate_homography else ())
        self._shape = None
        self._graphed = graphed
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def prepare(self, o: torch.Tensor, t: torch.Tensor):
        """Allocate the device buffers and (use_graph) warm up + capture the step for batches shaped like o / t.  Called on
        first use; ShardedPairPipeline calls it device by device from one thread, because a CUDA graph capture does not
        tolerate concurrent allocations from other threads."""
        if self._shape is not None:
            return
        with torch.cuda.device(self.device):
            self._prepare(o, t)

    def _prepare(self, o: torch.Tensor, t: torch.Tensor):
        dev = self.device
        self._shape = tuple(o.shape)
        self._copy = torch.cuda.Stream(device=dev)
        self._bufs = [(torch.empty(o.shape, dtype=o.dtype, device=dev), torch.empty(t.shape, dtype=t.dtype, device=dev))
                      for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._host_out = None
        if self.use_graph and self._graphed is None:
            self._bufs[0][0].copy_(o)
            self._bufs[0][1].copy_(t)
            self._graphed = self.pipe.capture(self._bufs[0][0], self._bufs[0][1])

    def _upload(self, i: int, o: torch.Tensor, t: torch.Tensor):
        do, dt = self._bufs[i % 2]
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(self._consumed[i % 2])          # the step that last read this buffer pair has finished
            do.copy_(o, non_blocking=True)
            dt.copy_(t, non_blocking=True)
            self._ready[i % 2].record(self._copy)
        self.h2d_bytes += o.numel() * o.element_size() + t.numel() * t.element_size()

    @torch.no_grad()
    def run(self, batches: Iterable, keep: bool = True) -> List[Dict[str, torch.Tensor]]:
        """keep=False returns only the last batch's results (timing loops that reuse one host batch)."""
        dev = self.device
        out: List[Dict[str, torch.Tensor]] = []
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            it = iter(batches)
            nxt = next(it, None)
            if nxt is None:
                return out
            if self._shape is None:
                self._prepare(*nxt)
            for ev in self._consumed:
                ev.record(main)
            self._upload(0, *nxt)
            pending = None
            i = 0
            while nxt is not None:
                if tuple(nxt[0].shape) != self._shape:
                    raise RuntimeError("PairStream: every batch must have the shape of the first one "
                                       f"({self._shape}); pad the last batch")
                cur, nxt = nxt, next(it, None)
                if nxt is not None:
                    self._upload(i + 1, *nxt)
                main.wait_event(self._ready[i % 2])
                o, t = self._bufs[i % 2]
                if self._graphed is not None:
                    self._graphed.load(o, t)                           # device-to-device into the graph's static inputs
                    self._consumed[i % 2].record(main)
                    r = self._graphed.replay()
                else:
                    r = self.pipe(o, t)
                    self._consumed[i % 2].record(main)
                vals = [getattr(r, k) for k in self.fields]
                if self._host_out is None:
                    self._host_out = [[torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in vals] for _ in range(2)]
                for h, x in zip(self._host_out[i % 2], vals):
                    h.copy_(x, non_blocking=True)
                    self.d2h_bytes += x.numel() * x.element_size()
                done = torch.cuda.Event()
                done.record(main)
                if pending is not None:
                    pending[0].synchronize()                           # results of step i-1 are on the host
                    if keep:
                        out.append({k: h.clone() for k, h in zip(self.fields, self._host_out[pending[1] % 2])})
                pending = (done, i)
                i += 1
            pending[0].synchronize()
            last = {k: h.clone() for k, h in zip(self.fields, self._host_out[pending[1] % 2])}
            if keep:
                out.append(last)
            else:
                out = [last]
        return out


def _concat(results: List[Dict[str, torch.Tensor]], n_valid: int) -> Dict[str, torch.Tensor]:
    if not results:
        return {}
    return {k: torch.cat([r[k] for r in results], 0)[:n_valid] for k in results[0]}


def _batches(optical: torch.Tensor, thermal: torch.Tensor, batch: int):
    """Fixed-shape batches; the last one is padded by repeating its final pair (trimmed again after the run)."""
    n = optical.shape[0]
    for lo in range(0, n, batch):
        o, t = optical[lo:lo + batch], thermal[lo:lo + batch]
        if o.shape[0] < batch:
            pad = batch - o.shape[0]
            o = torch.cat([o, o[-1:].expand(pad, *o.shape[1:])], 0)
            t = torch.cat([t, t[-1:].expand(pad, *t.shape[1:])], 0)
        yield o, t


class ShardedPairPipeline:
    """Batch-shards a host list of image pairs over GPUs (weights replicated, no collective on the data path).

    In-process mode (default): ``devices`` (all visible GPUs if None); one worker thread per device with its own weight
    replica, PairStream and CUDA graph.  Distributed mode (torch.distributed initialised and ``distributed=True``): this rank
    processes its ``shard_range`` on its own device and ``run`` gathers the per-rank results on every rank through the
    process group (gloo or nccl -- a host-side gather of keypoints / matches after the kernels, not part of the hot path).

    ``stream_factory(device) -> object with .run(batches) -> list of dicts`` replaces the CUDA worker (tests)."""

    def __init__(self, net=None, devices: Optional[Sequence] = None, batch: int = 64, distributed: bool = False,
                 stream_factory: Optional[Callable] = None, use_graph: bool = True, **pipe_kwargs):
        self.batch, self.distributed = batch, distributed
        self.pipe_kwargs, self.use_graph = pipe_kwargs, use_graph
        self.net = net
        if stream_factory is None:
            if devices is None:
                n = torch.cuda.device_count()
                if n == 0:
                    raise RuntimeError("xpoint_b200: no CUDA device; the pair pipeline has no CPU path")
                devices = [torch.device("cuda", torch.cuda.current_device())] if distributed else [torch.device("cuda", i) for i in range(n)]
            stream_factory = self._cuda_stream
        self.devices = list(devices) if devices is not None else [None]
        self._factory = stream_factory
        self._streams: Dict = {}

    def _cuda_stream(self, device):
        net = self.net if device == next(self.net.parameters()).device else copy.deepcopy(self.net).to(device)
        pipe = PairPipeline(net.eval(), **self.pipe_kwargs)
        return PairStream(pipe, device, use_graph=self.use_graph)

    def _stream(self, device):
        if device not in self._streams:
            self._streams[device] = self._factory(device)
        return self._streams[device]

    def _run_local(self, optical, thermal):
        """This process's pairs over its devices (threads); results concatenated in input order."""
        n = optical.shape[0]
        D = len(self.devices)
        parts: List[Optional[Dict]] = [None] * D
        errors: List[Optional[BaseException]] = [None] * D
        for i in range(D):                      # buffers / graph capture: one device after the other, on this thread
            lo, hi = shard_range(n, i, D)
            st = self._stream(self.devices[i])
            if hi > lo and hasattr(st, "prepare"):
                st.prepare(*next(_batches(optical[lo:hi], thermal[lo:hi], self.batch)))

        def work(i):
            try:
                lo, hi = shard_range(n, i, D)
                if hi > lo:
                    res = self._stream(self.devices[i]).run(_batches(optical[lo:hi], thermal[lo:hi], self.batch))
                    parts[i] = _concat(res, hi - lo)
            except BaseException as e:  # noqa: BLE001 (re-raised on the caller's thread)
                errors[i] = e
        if D == 1:
            work(0)
        else:
            threads = [threading.Thread(target=work, args=(i,)) for i in range(D)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
        for e in errors:
            if e is not None:
                raise e
        parts = [p for p in parts if p]
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]} if parts else {}

    def run(self, optical: torch.Tensor, thermal: torch.Tensor) -> Dict[str, torch.Tensor]:
        """optical, thermal: HOST tensors (N, 1, H, W) (pin them for asynchronous uploads).  Returns host tensors for all N
        pairs in input order (on every rank in distributed mode)."""
        if optical.shape != thermal.shape or optical.dim() != 4:
            raise RuntimeError("ShardedPairPipeline.run expects two (N, 1, H, W) host tensors of equal shape")
        import torch.distributed as dist
        if self.distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            rank, world = dist.get_rank(), dist.get_world_size()
            lo, hi = shard_range(optical.shape[0], rank, world)
            mine = self._run_local(optical[lo:hi], thermal[lo:hi]) if hi > lo else {}
            gathered: List = [None] * world
            dist.all_gather_object(gathered, {k: v.numpy() for k, v in mine.items()})
            keys = next((list(g.keys()) for g in gathered if g), [])
            return {k: torch.cat([torch.from_numpy(g[k]) for g in gathered if g], 0) for k in keys}
        return self._run_local(optical, thermal)
