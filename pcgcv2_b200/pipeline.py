"""Several frames in flight on one GPU.

One ``Codec.encode`` + ``Codec.decode`` leaves the GPU idle while the host runs the range coder (a strictly sequential
loop over the ~110 k bottleneck symbols, coder.py:46-70 / torchac) and while it waits at the two synchronising reads of
a frame.  Frames are independent (coder.py is batch-1 by construction, SURVEY.md section 8 e), so a ``FramePipeline``
keeps ``depth`` of them in flight: ``depth`` workers, each a host thread with its own ``Codec`` instance (own staging
buffers and overflow flag; the weights are read-only) and its own CUDA stream.  While one worker range-codes or waits,
the others enqueue and run kernels (ctypes calls and CUDA synchronisation release the GIL).  Results are identical to
the single-frame path frame by frame -- the same kernels run in the same order per frame.
"""
from __future__ import annotations

import queue
import threading

import torch

from .codec import Codec


class FramePipeline:
    def __init__(self, state_dict, device="cuda", depth: int = 2, **codec_kw):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("FramePipeline runs on a CUDA device only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth = depth
        self.codecs = [Codec(state_dict, device=self.device, **codec_kw) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self._jobs = [queue.SimpleQueue() for _ in range(depth)]
        self._done = queue.SimpleQueue()
        self._threads = [threading.Thread(target=self._worker, args=(w,), daemon=True, name=f"pcgc-frame-{w}")
                         for w in range(depth)]
        for t in self._threads:
            t.start()

    # ---------------------------------------------------------------- workers
    def _worker(self, w: int):
        codec, stream = self.codecs[w], self.streams[w]
        while True:
            job = self._jobs[w].get()
            if job is None:
                return
            idx, frame, rho, to_host, copy = job
            try:
                torch.cuda.set_device(self.device)
                with torch.cuda.stream(stream):
                    st = codec.encode(frame)
                    out = codec.decode(st, rho=rho, to_host=to_host)
                    if to_host and copy:
                        out = out.copy()                 # the codec's pinned staging buffer is reused by its next frame
                self._done.put((idx, st, out, None))
            except BaseException as e:                   # surfaced by roundtrip()
                self._done.put((idx, None, None, e))

    # ---------------------------------------------------------------- API
    def roundtrip(self, frames, rho: float = 1.0, to_host: bool = True, copy: bool = True):
        """encode + decode every frame (int32 [N,3] host arrays / pinned tensors / device tensors) ->
        list of (Stream, decoded coordinates) in input order.  Frame i runs on worker i % depth; the caller's current
        stream is ordered before the first and after the last kernel of the call."""
        frames = list(frames)
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)                           # inputs produced on the caller's stream
        for i, f in enumerate(frames):
            self._jobs[i % self.depth].put((i, f, rho, to_host, copy))
        results, err = [None] * len(frames), None
        for _ in frames:
            while True:
                try:
                    idx, st, out, e = self._done.get(timeout=1.0)
                    break
                except queue.Empty:
                    if not all(t.is_alive() for t in self._threads):
                        raise RuntimeError("FramePipeline: a worker thread died") from None
            err = err or e
            results[idx] = (st, out)
        for s in self.streams:
            cur.wait_stream(s)
        for r in results:                                # device results were allocated on a worker's stream: tell the caching
            if r is not None and isinstance(r[1], torch.Tensor) and r[1].is_cuda:   # allocator the caller's stream uses them too
                r[1].record_stream(cur)
        if err is not None:
            raise err
        return results

    def close(self):
        for q in self._jobs:
            q.put(None)
        for t in self._threads:
            t.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
