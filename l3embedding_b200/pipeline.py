"""Input pipeline of the B200 fit loop: AVC batch files -> pinned host buffers, ahead of the step.

The reference's `data_generator` (l3embedding/train.py:142-195) opens one gzip-HDF5 batch file after another, slices
and `np.concatenate`s them into batches and converts to float -- all in the training thread, serial with the step
(`use_multiprocessing` is commented out, train.py:411).  Here the same batch SEQUENCE is produced by two pieces:

* `BatchPlan`: the sequence as an index computation.  The reference's stream is "files in cyclic order (listing order
  first, `random.shuffle` after every full pass, seeded once with `random_state`), samples in file order, cut into
  consecutive runs of `batch_size`" -- so batch k is a short list of (file, start, stop) segments that can be
  computed from the per-file sample counts alone.  Resuming at `start_batch_idx` (train.py:164-169,184) is a seek in
  that plan: no data is read for the skipped batches.
* `PinnedBatchReader`: a reader thread that walks the plan and copies every segment STRAIGHT into a ring of
  page-locked uint8 / int16 / float32 buffers (no intermediate concatenation, no float conversion: `2*(u8/255)-1` and
  `pcm2float` run on the device, the H2D copy is 4x smaller).  The consumer gets zero-copy numpy views of a ring slot;
  a slot is not reused before two newer batches have been handed out, which covers the batch being computed and the
  one whose asynchronous upload (`l3_upload_batch_host`) is in flight.

File formats: the reference's HDF5 blobs (data/avc/sample.py:373-377,565-568: datasets `audio` (n,1,48000) int16,
`video` (n,224,224,3) uint8, `label` (n,2), gzip) through h5py when importable, else the built-in reader
(minihdf5.py); `.npz` files with the same keys.
"""
from __future__ import annotations

import os
import queue
import random
import threading
from collections import OrderedDict
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

KEYS = ("audio", "video", "label")
_SAMPLE_SHAPE = {"audio": (1, 48000), "video": (224, 224, 3), "label": (2,)}
_SAMPLE_DTYPE = {"audio": np.int16, "video": np.uint8, "label": np.float32}

Segment = Tuple[str, int, int]     # (path, first sample, one past the last sample)


# ---- batch files ---------------------------------------------------------------------------------------------------

class _Blob:
    """One batch file.  `count` needs metadata only; `read(key)` returns the whole dataset (decoded once, cached by the
    owner)."""

    def __init__(self, path: str):
        self.path = path
        self._np = self._h5 = self._mini = None
        if path.endswith(".npz"):
            self._np = np.load(path)
        else:
            try:
                import h5py   # the reference's reader, when it exists
                self._h5 = h5py.File(path, "r")
            except ImportError:
                from . import minihdf5
                self._mini = minihdf5.File(path)

    @property
    def count(self) -> int:
        if self._np is not None:
            # .npy header of the member: shape without decompressing the payload
            with self._np.zip.open("label.npy") as f:
                version = np.lib.format.read_magic(f)
                shape = (np.lib.format.read_array_header_1_0(f) if version == (1, 0)
                         else np.lib.format.read_array_header_2_0(f))[0]
            return int(shape[0])
        src = self._h5 if self._h5 is not None else self._mini
        return int(src["label"].shape[0])

    def read(self, key: str) -> np.ndarray:
        if self._np is not None:
            return self._np[key]
        src = self._h5 if self._h5 is not None else self._mini
        return np.asarray(src[key])

    def close(self):
        for h in (self._np, self._h5):
            if h is not None:
                h.close()


class BlobCache:
    """Decoded datasets of the most recently used batch files (a 1024-sample file is ~250 MB decoded)."""

    def __init__(self, max_files: int = 2):
        self.max_files = max_files
        self._data: "OrderedDict[str, Dict[str, np.ndarray]]" = OrderedDict()

    def get(self, path: str, key: str) -> np.ndarray:
        d = self._data.get(path)
        if d is None:
            d = {}
            self._data[path] = d
            while len(self._data) > self.max_files:
                self._data.popitem(last=False)
        else:
            self._data.move_to_end(path)
        if key not in d:
            b = _Blob(path)
            try:
                d[key] = b.read(key)
            finally:
                b.close()
        return d[key]


def blob_sample_count(path: str) -> int:
    b = _Blob(path)
    try:
        return b.count
    finally:
        b.close()


# ---- the plan ------------------------------------------------------------------------------------------------------

class BatchPlan:
    """The batch sequence of train.py:142-195 for one data directory, as (file, start, stop) segments per batch.

    File order: the directory listing, sorted (the reference iterates `os.listdir` in file-system order, which is not
    defined; sorted is one valid instance of it and makes runs reproducible), then -- as `cycle_shuffle` does --
    reshuffled in place after every full pass by ONE generator seeded with `random_state` (train.py:134-144)."""

    def __init__(self, data_dir: str, batch_size: int, random_state: int = 20180123,
                 files: Optional[Sequence[str]] = None):
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        self.data_dir = data_dir
        self.batch_size = int(batch_size)
        self.random_state = random_state
        names = sorted(os.listdir(data_dir)) if files is None else list(files)
        if not names:
            raise ValueError("no batch files in %s" % data_dir)
        self.paths = [os.path.join(data_dir, n) for n in names]
        self._counts: Dict[str, int] = {}

    def count(self, path: str) -> int:
        c = self._counts.get(path)
        if c is None:
            c = self._counts[path] = blob_sample_count(path)
        return c

    def _file_stream(self) -> Iterator[str]:
        rng = random.Random(self.random_state)
        order = list(self.paths)
        while True:
            yield from list(order)
            rng.shuffle(order)

    def segments(self, start_batch: int = 0) -> Iterator[List[Segment]]:
        """Yields the segment list of batch `start_batch`, `start_batch + 1`, ...  Seeking costs one metadata look-up per
        file passed over, never a data read."""
        to_skip = int(start_batch) * self.batch_size      # samples before the first wanted batch
        need = self.batch_size
        cur: List[Segment] = []
        for path in self._file_stream():
            n = self.count(path)
            pos = 0
            if to_skip >= n:
                to_skip -= n
                continue
            pos, to_skip = to_skip, 0
            while pos < n:
                take = min(need, n - pos)
                cur.append((path, pos, pos + take))
                pos += take
                need -= take
                if need == 0:
                    yield cur
                    cur, need = [], self.batch_size


# ---- pinned ring + reader thread --------------------------------------------------------------------------------

def _alloc(shape, dtype, pinned: bool):
    """(array, owner): a page-locked buffer when CUDA is there (torch allocates it; `owner` keeps it alive), else plain
    host memory."""
    if pinned:
        try:
            import torch
            if torch.cuda.is_available():
                tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
                       np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
                t = torch.empty(tuple(shape), dtype=tdt).pin_memory()
                return t.numpy(), t
        except ImportError:
            pass
    return np.empty(shape, dtype), None


class PinnedBatchReader:
    """Iterator over the batches of a BatchPlan, filled by a background thread into a ring of pinned host buffers.

    Every item is a dict {'video': (B,224,224,3) uint8, 'audio': (B,1,48000) int16, 'label': (B,2) float32} of views
    into one ring slot; an item stays valid until TWO further items have been taken from the iterator."""

    def __init__(self, plan: BatchPlan, start_batch: int = 0, max_batches: Optional[int] = None, slots: int = 4,
                 keys: Sequence[str] = KEYS, pinned: bool = True, cache_files: int = 2):
        if slots < 3:
            raise ValueError("the ring needs at least 3 slots (one being computed, one being uploaded, one being filled)")
        self.plan, self.keys, self.slots = plan, tuple(keys), slots
        B = plan.batch_size
        self._owners = []
        self._ring = []
        for _ in range(slots):
            slot = {}
            for k in self.keys:
                slot[k], owner = _alloc((B,) + _SAMPLE_SHAPE[k], _SAMPLE_DTYPE[k], pinned)
                self._owners.append(owner)
            self._ring.append(slot)
        # `_ready` bounds how far the reader runs ahead (a place is reserved before a slot is filled and freed when the
        # consumer takes the batch); `_filled` hands the filled slots over in order
        self._ready: "queue.Queue" = queue.Queue(maxsize=slots - 2)
        self._lock = threading.Condition()
        self._filled: list = []
        self._stop = threading.Event()
        self._cache = BlobCache(cache_files)
        self._start, self._max = int(start_batch), max_batches
        self._thread = threading.Thread(target=self._run, name="l3-batch-reader", daemon=True)
        self._thread.start()

    def _fill(self, slot: Dict[str, np.ndarray], segs: List[Segment]):
        at = 0
        for path, s, e in segs:
            for k in self.keys:
                src = self._cache.get(path, k)
                dst = slot[k][at:at + (e - s)]
                if k == "label":
                    np.copyto(dst, np.asarray(src[s:e]).reshape(dst.shape), casting="unsafe")
                else:
                    np.copyto(dst, src[s:e].reshape(dst.shape))
            at += e - s

    def _run(self):
        try:
            for j, segs in enumerate(self.plan.segments(self._start)):
                if self._stop.is_set() or (self._max is not None and j >= self._max):
                    break
                slot = self._ring[j % self.slots]
                # a free place in the queue means batch j - (slots - 2) has been handed out; slot j % slots last held
                # batch j - slots, i.e. at least two newer batches have been taken since
                while not self._stop.is_set():
                    try:
                        self._ready.put(None, timeout=0.1)      # reserve the place first (blocks when the ring is full)
                        break
                    except queue.Full:
                        continue
                if self._stop.is_set():
                    break
                self._fill(slot, segs)
                with self._lock:
                    self._filled.append(slot)
                    self._lock.notify()
        except BaseException as e:     # surfaced in the consumer
            with self._lock:
                self._filled.append(e)
                self._lock.notify()
        else:
            with self._lock:
                self._filled.append(StopIteration())
                self._lock.notify()

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, np.ndarray]:
        with self._lock:
            while not self._filled:
                self._lock.wait()
            item = self._filled.pop(0)
        if isinstance(item, StopIteration):
            with self._lock:
                self._filled.append(item)
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        self._ready.get()            # frees one place: the reader may fill the next slot
        return item

    def close(self):
        self._stop.set()
        try:
            while True:
                self._ready.get_nowait()
        except queue.Empty:
            pass
        self._thread.join(timeout=5)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_batches(plan: BatchPlan, start_batch: int = 0, keys: Sequence[str] = KEYS) -> Iterator[Dict[str, np.ndarray]]:
    """The same batches without a thread or a ring (fresh arrays per batch): the synchronous path of the tests."""
    cache = BlobCache(2)
    for segs in plan.segments(start_batch):
        out = {}
        for k in keys:
            parts = [cache.get(p, k)[s:e] for p, s, e in segs]
            a = parts[0] if len(parts) == 1 else np.concatenate(parts)
            out[k] = np.ascontiguousarray(a, dtype=_SAMPLE_DTYPE[k] if k in _SAMPLE_DTYPE else None)
        yield out
