"""Device-buffer plumbing (torch is used for memory, streams and H2D/D2H copies only)."""
import numpy as np
import torch

_DT = {np.dtype('float32'): torch.float32, np.dtype('int32'): torch.int32, np.dtype('uint32'): torch.uint32,
       np.dtype('uint8'): torch.uint8}


def cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError('fbs_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def is_host(x) -> bool:
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def dev(x, dtype, device=None):
    """Contiguous device tensor of ``dtype`` (torch dtype).  numpy / python / CPU tensors are copied H2D."""
    device = device or cuda_device()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        np_dt = {torch.float32: np.float32, torch.int32: np.int32, torch.uint32: np.uint32, torch.uint8: np.uint8}[dtype]
        t = torch.from_numpy(np.ascontiguousarray(a.astype(np_dt, copy=False)))
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        if t.is_pinned():
            t = t.to(device, non_blocking=True)
        else:
            t = t.to(device)
    return t.contiguous()


def empty(shape, dtype, device=None):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=device or cuda_device())


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def out(t, host: bool):
    """Return a device tensor as-is, or as a numpy array when the caller passed host buffers.

    The D2H copy lands in page-locked memory from torch's caching host allocator (a pageable destination makes the
    driver stage the copy through a bounce buffer at a fraction of the PCIe rate); the numpy array returned is a view
    of that block and keeps it alive, and feeding it back as an input is again a pinned H2D copy.
    """
    if t is None or not host:
        return t
    if t.dtype == torch.uint32 or t.numel() < 4096:
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()
