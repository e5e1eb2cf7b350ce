"""HDF5 -> device input pipeline for real training data.

Reference path: ``OpenFOAMDataRepository.read_data`` (turbdiff/data/ofles.py:396-418) reads, per variable, the rows
``dataset[unique_sorted_idxs]`` of a channels-last ``(T, n_cells, F)`` HDF5 dataset (h5py wants sorted unique indices),
undoes the sorting with the inverse index, and hands the tensors to ``OpenFOAMData.grid_embedding`` (:220-240) and
``Normalization.normalize_grid`` (models/normalization.py:20-24) - three host passes, one pageable H2D copy and four
eager device passes per batch.

Here the same read lands directly in a pinned staging buffer laid out as the kernel wants it (``(B, n_cells, F)`` with all
variables concatenated on the channel axis), travels on a copy stream, and ONE launch of ``tdb_scatter_normalize`` builds the
normalised, boundary-filled ``(B, F, X, Y, Z)`` grid on the device.  ``depth`` staging slots are kept in flight by a worker
thread, so batch k+1 is read and copied while batch k trains.

The datasets are duck-typed (``ds[indices] -> array (len, n_cells[, F])``, optionally h5py's ``read_direct``): h5py is not
part of this image, so the tests drive the pipeline with numpy arrays / memmaps; an ``h5py.Dataset`` satisfies the same
protocol.  Results are bit-identical to the reference's op sequence (tests/test_gpu_pipeline.py)."""

from __future__ import annotations

import queue
import threading
from dataclasses import dataclass

import numpy as np
import torch

from .models import utils as U


def sorted_unique_read_plan(sample_idxs):
    """(unique_sorted_idxs, inverse_idx) exactly as ``read_data`` derives them (ofles.py:398-399): the rows are read once,
    in file order, and ``rows[inverse_idx]`` restores the requested order (duplicates included)."""
    idx = np.asarray(sample_idxs)
    return np.unique(idx, return_inverse=True)


def read_channels_last(datasets, sample_idxs, out: np.ndarray | None = None) -> np.ndarray:
    """``read_data`` for all variables at once: ``datasets`` is an ordered list of ``(dataset, dims)``; returns (and fills
    ``out`` when given) a ``(B, n_cells, sum(dims))`` float32 array - scalar fields get their feature axis (ofles.py:410-411),
    the sorting / uniquification is undone (:414)."""
    uniq, inv = sorted_unique_read_plan(sample_idxs)
    B = len(inv)
    F = sum(d for _, d in datasets)
    f0 = 0
    for ds, dims in datasets:
        rows = np.asarray(ds[uniq])  # h5py: one hyperslab read of the sorted unique rows
        if rows.ndim == 2:
            rows = rows[..., None]
        if rows.shape[-1] != dims:
            raise ValueError(f"turbdiff_b200.pipeline: dataset has {rows.shape[-1]} features, expected {dims}")
        if out is None:
            out = np.empty((B, rows.shape[1], F), dtype=np.float32)
        out[:B, :, f0 : f0 + dims] = rows[inv]
        f0 += dims
    return out[:B]


@dataclass
class DeviceBatch:
    """One batch on the device: ``x`` = normalised grid (B, F, X, Y, Z) ready for ``GaussianDiffusion.forward``; ``samples``
    = the channels-last cell values (B, n_cells, F) as read; ``idxs`` = the requested sample indices."""

    x: torch.Tensor
    samples: torch.Tensor
    idxs: np.ndarray
    ready: torch.cuda.Event

    def wait(self, stream: torch.cuda.Stream | None = None):
        (stream or torch.cuda.current_stream()).wait_event(self.ready)
        return self


class DeviceBatchPipeline:
    """Prefetching reader: ``for batch in DeviceBatchPipeline(...).run(index_batches): loss = model(batch.x, ...)``.

    datasets     ordered list of ``(dataset, dims)`` per variable (reference: ``data_group[v.name.lower()]``, ``v.dims``)
    cell_idx     int64 flat voxel index of every mesh cell (``grid/cell_idx``)
    cell_counts  padded grid shape (X, Y, Z)
    mean, std    per-channel normalisers (``OpenFOAMStats.normalizers``)
    fixed_values FIXED_VALUE boundary writes, see ``models.utils.fixed_values_of``
    """

    def __init__(self, datasets, cell_idx, cell_counts, mean, std, fixed_values=None, device=None, max_batch=8, depth=2):
        self.datasets = list(datasets)
        self.F = sum(d for _, d in self.datasets)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("turbdiff_b200.pipeline: the device side of the pipeline needs a CUDA device (no CPU path)")
        self.cell_counts = tuple(int(c) for c in cell_counts)
        self.cell_idx = torch.as_tensor(cell_idx, dtype=torch.int64).to(self.device)
        self.n_cells = int(self.cell_idx.numel())
        self.mean, self.std = torch.as_tensor(mean, dtype=torch.float32), torch.as_tensor(std, dtype=torch.float32)
        nvox = self.cell_counts[0] * self.cell_counts[1] * self.cell_counts[2]
        # per-geometry tables of the scatter kernel: built once
        self.tables = U.boundary_code(self.cell_idx, nvox, self.F, fixed_values)
        self.depth = depth
        self.max_batch = max_batch
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [torch.empty((max_batch, self.n_cells, self.F), dtype=torch.float32, pin_memory=True) for _ in range(depth)]
        self._slot_free = [None] * depth  # event: the H2D copy out of the slot has completed

    def _produce(self, slot: int, idxs) -> DeviceBatch:
        idxs = np.asarray(idxs)
        B = len(idxs)
        if B > self.max_batch:
            raise ValueError(f"turbdiff_b200.pipeline: batch of {B} exceeds max_batch={self.max_batch}")
        if self._slot_free[slot] is not None:
            self._slot_free[slot].synchronize()  # the previous copy out of this staging buffer is done
        host = self._slots[slot]
        read_channels_last(self.datasets, idxs, out=host.numpy())
        with torch.cuda.stream(self.copy_stream):
            dev = host[:B].to(self.device, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.copy_stream)
            x = U.scatter_normalize(dev, self.cell_idx, self.cell_counts, self.mean, self.std, tables=self.tables)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        self._slot_free[slot] = copied
        return DeviceBatch(x, dev, idxs, ready)

    def load(self, idxs) -> DeviceBatch:
        """Synchronous form: one batch, usable on the current stream when it returns."""
        return self._produce(0, idxs).wait()

    def run(self, index_batches):
        """Generator over device batches with ``depth`` batches in flight (reads + copies on a worker thread)."""
        q: queue.Queue = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()

        def worker():
            torch.cuda.set_device(dev_index)
            try:
                for k, idxs in enumerate(index_batches):
                    if stop.is_set():
                        break
                    q.put(self._produce(k % self.depth, idxs))
                q.put(None)
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        th = threading.Thread(target=worker, daemon=True, name="tdb-pipeline")
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                item.wait()
                # the consumer's stream owns the tensors from here on
                item.x.record_stream(torch.cuda.current_stream())
                item.samples.record_stream(torch.cuda.current_stream())
                yield item
        finally:
            stop.set()
            while th.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.05)
