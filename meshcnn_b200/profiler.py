"""CUDA-event timers around individual C-ABI calls (bench.py's live roofline measurement).  Off by default; when on,
every call site records an event pair on the current stream -- no synchronisation until ``collect()``."""
import contextlib

import torch

enabled = False
_records = []


@contextlib.contextmanager
def span(name, **meta):
    if not enabled:
        yield
        return
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _records.append((name, meta, a, b))


def reset():
    del _records[:]


def collect():
    """-> list of (name, meta, milliseconds); synchronises."""
    torch.cuda.synchronize()
    out = [(n, m, a.elapsed_time(b)) for n, m, a, b in _records]
    reset()
    return out
