"""Input staging for the training loop. The reference converts every batch synchronously
torch -> numpy -> jnp on the step's critical path (trainer.py:258, 546-547); here host batches are pinned and copied
to the device on a dedicated copy stream `depth` steps ahead, so the host-to-device transfer of step i+1 overlaps the
kernels of step i. Values are passed through untouched (the reference does not rescale MovingMNIST either).

Measured on config_v2_2 (655 KB per batch): the plain `TrainStep.step(pinned_host_batch, ...)` path, which copies
into the step's resident input buffer on the compute stream, is FASTER end to end (582 vs 562 clips/s) - at this
batch size the transfer is 30 us and the prefetcher's per-batch allocation / event bookkeeping costs more than it
hides. `bench.py` therefore uses the plain path; the prefetcher is for large clips (v2_3x: 4 MB per batch and up)."""
from __future__ import annotations

from collections import deque
from typing import Iterable, Iterator

import torch


class DevicePrefetcher:
    """Iterates device-resident batches of an iterable of host tensors (b, c, f, h, w), `depth` batches in flight."""

    def __init__(self, batches: Iterable[torch.Tensor], device="cuda", depth: int = 2):
        self.it: Iterator[torch.Tensor] = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue: deque = deque()
        for _ in range(max(1, depth)):
            self._enqueue()

    def _enqueue(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            return
        if not host.is_pinned():
            host = host.contiguous().pin_memory()
        with torch.cuda.stream(self.stream):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev, host))  # the pinned source stays alive until its copy has been consumed

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        if not self.queue:
            raise StopIteration
        dev, ev, _host = self.queue.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        dev.record_stream(cur)
        self._enqueue()
        return dev
