"""Multi-GPU plumbing: one process per GPU, streams sharded across ranks, and the
only collective on the path -- the broadcast of the derived plan
(fsk_b200_rx_params, plain bytes) from rank 0, so that every rank runs the exact
integers rank 0 derived.  NCCL on GPUs, gloo in the CPU tests."""
import ctypes as C

from .api import RxParams


def broadcast_params(params, src=0, device=None):
    """params: RxParams on `src` (ignored elsewhere).  Returns the RxParams every rank
    should build its engine from."""
    import torch
    import torch.distributed as dist
    n = C.sizeof(RxParams)
    blob = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        blob.copy_(torch.frombuffer(bytearray(bytes(params)), dtype=torch.uint8))
    dist.broadcast(blob, src=src)
    return RxParams.from_buffer_copy(blob.cpu().numpy().tobytes())


def shard_range(nstreams, rank, world):
    """Contiguous block of streams owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(nstreams), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
