"""Process-wide default engine for stand-alone device calls (single polynomial solves, collision
queries outside a planner).  Planners own their private ``Engine``."""
import threading

_lock = threading.Lock()
_default = {}


def current_device_and_stream():
    """torch supplies the device index and the stream handle (PyTorch = plumbing only)."""
    import torch
    if not torch.cuda.is_available():
        from commonroad_rp_b200._lib import RpError
        raise RpError("no CUDA device available: the candidate-trajectory path has no CPU fallback")
    dev = torch.cuda.current_device()
    return dev, torch.cuda.current_stream(dev).cuda_stream


def default_engine():
    from commonroad_rp_b200._lib import Engine
    dev, stream = current_device_and_stream()
    with _lock:
        key = (dev, stream)
        if key not in _default:
            _default[key] = Engine(dev, stream)
        return _default[key]
