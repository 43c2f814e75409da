"""Input pipelines in front of the train step.

The reference feeds the step from ``DataLoader(dataset, batch_size, shuffle=True, pin_memory=True)``
(`train.py:153-158`) and copies every batch host->device inside `train_loop`
(`BGT/models/train.py:118-123`).  At B200 step rates both become the bottleneck, so two replacements
are offered:

* ``HostBatchPrefetcher`` — batches stay in (pinned) HOST memory; the H2D copy of batch i+1 runs on a
  copy stream while the compute stream runs step i (two device buffer pairs, event hand-off in both
  directions).  This is the path `bench.py` times as ``e2e``.
* ``DeviceResidentLoader`` — the whole ``processed_inputs / processed_outputs`` pair
  (`dataset.py:352-356`) lives on the GPU; every epoch draws one permutation and each batch is one
  row gather into a reused buffer: no per-sample ``__getitem__``, no collate, no H2D traffic.  It yields
  ``(x, y, idx)`` like the reference's dataset (`dataset.py:352-356`) and exposes ``.dataset`` with a
  length, so it can be passed to ``train_loop`` as the ``dataloader`` argument unchanged.

PyTorch is used for memory, streams and events only; no arithmetic happens here.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch


class HostBatchPrefetcher:
    """Double-buffered host->device feeder.

    ``submit(xh, yh)`` enqueues the copy of one host batch into the free device slot on the copy stream
    (after the compute stream has finished with that slot's previous contents); ``get()`` makes the
    current stream wait for the oldest submitted copy and returns its device tensors.  At most two
    batches may be in flight."""

    def __init__(self, device, x_shape, y_shape, dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostBatchPrefetcher feeds a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [(torch.empty(x_shape, dtype=dtype, device=self.device), torch.empty(y_shape, dtype=dtype, device=self.device))
                      for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]      # copy of slot i finished (recorded on the copy stream)
        self.released = [None, None]                               # compute finished reading slot i (recorded on the compute stream)
        self._head = 0                                             # next slot to fill
        self._tail = 0                                             # next slot to hand out
        self._inflight = 0
        self._held: Optional[int] = None
        self.h2d_bytes = 0

    def submit(self, xh: torch.Tensor, yh: torch.Tensor) -> None:
        if self._inflight >= 2:
            raise RuntimeError("HostBatchPrefetcher: both device slots are in flight; call get() first")
        i = self._head
        x, y = self.slots[i]
        with torch.cuda.stream(self.copy_stream):
            if self.released[i] is not None:
                self.copy_stream.wait_event(self.released[i])
            x.copy_(xh, non_blocking=True)
            y.copy_(yh, non_blocking=True)
            self.copied[i].record(self.copy_stream)
        self.h2d_bytes += xh.numel() * xh.element_size() + yh.numel() * yh.element_size()
        self._head ^= 1
        self._inflight += 1

    def get(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device tensors of the oldest submitted batch, valid until the next-but-one ``submit``."""
        if self._inflight == 0:
            raise RuntimeError("HostBatchPrefetcher: nothing submitted")
        self.release()
        i = self._tail
        torch.cuda.current_stream(self.device).wait_event(self.copied[i])
        self._tail ^= 1
        self._inflight -= 1
        self._held = i
        return self.slots[i]

    def release(self) -> None:
        """Mark the batch handed out by the last ``get()`` as consumed by everything enqueued so far."""
        if self._held is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.released[self._held] = ev
            self._held = None


class DeviceResidentLoader:
    """Iterates ``(x, y, idx)`` batches of a dataset that lives on the GPU.

    inputs ``[S, 32, E_src]`` and outputs ``[S, 32, 27]`` are copied to the device once.  Each epoch
    (``__iter__``) draws a permutation with a device generator (``shuffle=True`` like the reference's
    loader) and every batch is a row gather into one of two reused buffers.  ``drop_last=False`` yields
    the ragged tail batch like torch's DataLoader default."""

    def __init__(self, inputs, outputs, batch_size: int, device, shuffle: bool = True, drop_last: bool = False, seed: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceResidentLoader keeps the dataset on a CUDA device")
        self.x = torch.as_tensor(inputs, dtype=torch.float32).to(self.device).contiguous()
        self.y = torch.as_tensor(outputs, dtype=torch.float32).to(self.device).contiguous()
        if self.x.shape[0] != self.y.shape[0]:
            raise ValueError("inputs and outputs hold a different number of sequences")
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.dataset = range(self.x.shape[0])            # train_loop only asks for len(dataloader.dataset)
        n = min(self.batch_size, self.x.shape[0])
        self._buf = [(torch.empty((n,) + tuple(self.x.shape[1:]), device=self.device), torch.empty((n,) + tuple(self.y.shape[1:]), device=self.device))
                     for _ in range(2)]

    def __len__(self) -> int:
        s = self.x.shape[0]
        return s // self.batch_size if self.drop_last else (s + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        s = self.x.shape[0]
        order = torch.randperm(s, device=self.device, generator=self.gen) if self.shuffle else torch.arange(s, device=self.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            bx, by = self._buf[b & 1]
            k = idx.numel()
            torch.index_select(self.x, 0, idx, out=bx[:k])
            torch.index_select(self.y, 0, idx, out=by[:k])
            yield bx[:k], by[:k], idx
