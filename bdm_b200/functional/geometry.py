"""Coordinate-only work of a denoiser forward, memoised and run ahead on a side stream.

Everything the sparse path derives from point *coordinates* alone -- the FPS pyramid, ball-query
neighbour lists, 3-NN (index, weight) pairs, voxel plans -- is independent of the features, i.e. of
every dense layer.  FPS in particular is a chain of ~1350 strictly dependent rounds that keeps 16 of
148 SMs busy for ~0.5 ms per forward.  `GeometryCache.plan_ahead` issues that whole chain on a second
CUDA stream at the top of the forward, where it overlaps with the first PVConv's 3-D convolutions;
the modules later find the results in the memo (keyed by the identity of their input tensors) and
only make the main stream wait on the producing event.  The arithmetic is the same kernels on the
same inputs: results are bit-identical to the inline order.

Active only for inference on CUDA (no autograd graph is built through the memoised tensors).
"""
import contextlib

import torch

_active = None  # the GeometryCache of the forward in flight, if any


def active():
    return _active


def tensor_version(t):
    """In-place edit counter of a tensor; inference tensors (torch.inference_mode) do not track one -- they
    cannot be edited in place outside inference mode either, and the memo holds a reference to the tensor
    object itself, so identity alone is a sound key there."""
    return 0 if t.is_inference() else t._version


def _tensor_key(t):
    return (t.data_ptr(), tensor_version(t), tuple(t.shape), tuple(t.stride()))


class _Entry:
    __slots__ = ("value", "event", "stream_id", "keep")

    def __init__(self, value, event, stream_id, keep):
        self.value, self.event, self.stream_id, self.keep = value, event, stream_id, keep


def _tensors_in(value):
    if isinstance(value, torch.Tensor):
        yield value
    elif isinstance(value, (tuple, list)):
        for v in value:
            yield from _tensors_in(v)
    elif hasattr(value, "__slots__"):
        for name in value.__slots__:
            yield from _tensors_in(getattr(value, name))


class GeometryCache:
    def __init__(self):
        self.entries = {}
        self.side = None

    def get(self, op, tensors, scalars, compute):
        """Memoised `compute()`; `tensors` are the inputs whose identity defines the key."""
        key = (op,) + tuple(_tensor_key(t) for t in tensors) + tuple(scalars)
        cur = torch.cuda.current_stream()
        ent = self.entries.get(key)
        if ent is None:
            value = compute()
            ev = torch.cuda.Event()
            ev.record(cur)
            # `keep` pins the input tensors: a freed-and-reused address must not alias a stale key
            ent = _Entry(value, ev, cur.cuda_stream, tuple(tensors))
            self.entries[key] = ent
            return value
        if ent.stream_id != cur.cuda_stream:
            cur.wait_event(ent.event)
            for t in _tensors_in(ent.value):
                t.record_stream(cur)
            ent.stream_id = cur.cuda_stream  # later uses on this stream are ordered already
        return ent.value

    @contextlib.contextmanager
    def side_stream(self):
        main = torch.cuda.current_stream()
        if self.side is None:
            self.side = torch.cuda.Stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            yield
        # no join here: consumers wait per entry.  `close()` joins before the forward returns.

    def close(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.entries.clear()


@contextlib.contextmanager
def scope(enabled=True):
    """Activate a GeometryCache for the duration of one forward pass."""
    global _active
    if not enabled or _active is not None:
        yield _active
        return
    _active = GeometryCache()
    try:
        yield _active
    finally:
        cache, _active = _active, None
        cache.close()


_last = {}  # op -> (key, pinned input tensors, value): one-entry memo used outside a scope


def memo(op, tensors, scalars, compute, keep_last=False):
    """Inside a scope: memoised by the identity of `tensors`.  Outside: recomputed, unless `keep_last`
    asks for a one-entry memo (consecutive calls on the same tensor objects, same stream)."""
    if _active is not None:
        return _active.get(op, tensors, scalars, compute)
    if not keep_last:
        return compute()
    key = tuple(_tensor_key(t) for t in tensors) + tuple(scalars) + (torch.cuda.current_stream().cuda_stream
                                                                     if tensors and tensors[0].is_cuda else 0,)
    hit = _last.get(op)
    if hit is not None and hit[0] == key and all(a is b for a, b in zip(hit[1], tensors)):
        return hit[2]
    value = compute()
    _last[op] = (key, tuple(tensors), value)
    return value
