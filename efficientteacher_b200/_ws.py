"""Caller-owned scratch buffers for the C ABI (the library never allocates)."""
import torch

_cache = {}


def workspace(name, nbytes, device):
    """A uint8 CUDA buffer of at least `nbytes`, cached per (name, device) and grown on demand."""
    key = (name, str(device))
    buf = _cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _cache[key] = buf
    return buf
