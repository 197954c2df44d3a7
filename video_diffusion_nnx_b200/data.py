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
from functools import partial
from typing import Iterable, Iterator, Optional

import numpy as np
import torch


def identity(t, *args, **kwargs):
    return t


def cast_num_frames(t: np.ndarray, *, frames: int) -> np.ndarray:
    """utils.py:380-397: truncate or zero-pad a (c, f, h, w) clip along the frame axis."""
    f = t.shape[1]
    if f == frames:
        return t
    if f > frames:
        return t[:, :frames, ...]
    return np.pad(t, ((0, 0), (0, frames - f), (0, 0), (0, 0)))


class MovingMNIST(torch.utils.data.Dataset):
    """datasets.py:10-64: Moving MNIST clips from a `.npy` file of shape (frames, sequences, h, w).

    Behaviour reproduced from the reference, quirks included: `__getitem__` returns the RAW float32
    array `(c=1, f, h, w)` truncated / zero-padded to `num_frames` - pixel values stay 0..255, and neither the resize
    nor the horizontal flip the constructor accepts is ever applied (the reference builds `self.transform`,
    datasets.py:50-55, and never calls it). Same constructor arguments and attribute names (including `channnels`)."""

    def __init__(self, file_path, image_size, channels: int = 1, num_frames: int = 20, horizontal_flip: bool = False,
                 force_num_frames: bool = True):
        super().__init__()
        self.file_path = file_path
        self.image_size = image_size
        self.channnels = channels  # sic (datasets.py:39)
        self.horizontal_flip = horizontal_flip
        arrays = np.load(file_path)  # (f, b, h, w)
        self.arrays = np.ascontiguousarray(np.transpose(arrays, (1, 0, 2, 3))[:, None, ...]).astype(np.float32)
        self.cast_num_frames_fn = partial(cast_num_frames, frames=num_frames) if force_num_frames else identity

    def __len__(self) -> int:
        return self.arrays.shape[0]

    def __getitem__(self, index) -> np.ndarray:
        return self.cast_num_frames_fn(self.arrays[index])


def cycle(dl):
    """utils.py:72-83: endless iteration over a data loader."""
    while True:
        for batch in dl:
            yield batch


def training_batches(dataset, batch_size: int, *, shuffle: bool = True, seed: int = 0, rank: int = 0, world: int = 1,
                     device: Optional[str] = "cuda", depth: int = 2):
    """The reference's input pipeline (trainer.py:247-258, 546-547) for one data-parallel rank: an endless, shuffled,
    drop-last stream of GLOBAL batches of `world * batch_size` clips of which this rank keeps its shard
    (trainer.py:307-309 shards the batch over the `data` axis), pinned on the host and - with `device` - staged onto the
    GPU `depth` batches ahead by DevicePrefetcher. Every rank must use the same `seed`."""
    g = torch.Generator().manual_seed(seed)
    dl = torch.utils.data.DataLoader(dataset, batch_size=world * batch_size, shuffle=shuffle, drop_last=True,
                                     pin_memory=False, generator=g)

    def shards():
        for batch in cycle(dl):
            yield torch.as_tensor(batch)[rank * batch_size:(rank + 1) * batch_size].contiguous()

    return DevicePrefetcher(shards(), device=device, depth=depth) if device is not None else shards()


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
