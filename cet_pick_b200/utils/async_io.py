"""Overlap of the host-side stages around the detector (north_star (3)): a reader thread that
prepares tomogram i+1 (file read + GPU pre-processing on its own CUDA stream) while tomogram i is in the detector,
and a small pool of writer threads that move finished heat-maps from page-locked staging buffers to `<name>_hm.mrc`
while the GPU is already on the next tomogram.  The reference overlaps only the read, with a one-worker DataLoader
(cet_pick/test.py:77); its `run()` blocks on the 268 MB heat-map copy and the file write (tomo_det.py:58-67)."""
from __future__ import annotations

import queue
import threading

import torch


class AsyncWriter:
    """N daemon threads executing submitted jobs in FIFO order; `flush()` waits for all of them and re-raises the first
    exception a job hit (a failed write must not pass silently)."""

    def __init__(self, threads: int = 3, max_pending: int = 6):
        self._q = queue.Queue(maxsize=max_pending)
        self._err = []
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, threads))]
        for t in self._threads:
            t.start()

    def _work(self):
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                job()
            except BaseException as e:       # noqa: BLE001 - reported by flush()
                self._err.append(e)
            finally:
                self._q.task_done()

    def submit(self, job):
        if self._err:
            raise self._err.pop(0)
        self._q.put(job)                     # blocks when max_pending jobs are queued: bounds the staging memory

    def flush(self):
        self._q.join()
        if self._err:
            raise self._err.pop(0)

    def close(self):
        self.flush()
        for _ in self._threads:
            self._q.put(None)


class PinnedPool:
    """page-locked staging buffers of one shape, handed out round-robin; a buffer is reusable once its job released it"""

    def __init__(self, count: int = 4):
        self._free = queue.Queue()
        self._count, self._made, self._key = count, 0, None

    def get(self, shape, dtype):
        key = (tuple(shape), dtype)
        if key != self._key:                 # new volume size: drop the old buffers (they are freed when released)
            # page-locking a 268 MB buffer takes ~0.1 s: all `count` buffers are made at the first use of a size (the
            # first tomogram of a run), not one by one whenever the writers fall behind
            self._key, self._made = key, self._count
            self._free = queue.Queue()
            for _ in range(self._count):
                self._free.put(torch.empty(shape, dtype=dtype, pin_memory=True))
        return self._free.get()

    def put(self, buf):
        if (tuple(buf.shape), buf.dtype) == self._key:
            self._free.put(buf)


class Prefetcher:
    """iterate `make(item)` over `items` with the NEXT result prepared by a background thread on its own CUDA stream.
    `make` returns any object; tensors inside it were produced on the side stream, so the consumer waits on the
    recorded event before touching them (done here: `__next__` returns after `event.synchronize()`-free stream wait)."""

    def __init__(self, items, make, device=None, depth: int = 1):
        self._items, self._make = list(items), make
        self._device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._q = queue.Queue(maxsize=max(1, depth))
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        try:
            torch.cuda.set_device(self._device)
            side = torch.cuda.Stream(self._device)
            for it in self._items:
                with torch.cuda.stream(side):
                    out = self._make(it)
                    ev = torch.cuda.Event()
                    ev.record(side)
                self._q.put((it, out, ev, None))
        except BaseException as e:           # noqa: BLE001 - re-raised in the consumer
            self._q.put((None, None, None, e))
        self._q.put(None)

    def __iter__(self):
        return self

    def __next__(self):
        r = self._q.get()
        if r is None:
            raise StopIteration
        it, out, ev, err = r
        if err is not None:
            raise err
        torch.cuda.current_stream(self._device).wait_event(ev)
        return it, out
