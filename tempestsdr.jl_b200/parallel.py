"""Multi-GPU host logic: one process per GPU.  On GPUs the combine runs behind the C ABI (tsdr_comm_* /
tsdr_chain_allreduce: NCCL bound by the library itself, the EMA tail weight folded into the collective);
torch.distributed only launches the ranks and carries the 128-byte communicator id.  The gloo path below
(allreduce_partial on CPU tensors) exists for the CPU tests of the host logic.

The chain shards without any data-path exchange:
  * independent recv! buffers          -> buffer b on rank b mod N          (cfg 3)
  * independent hypotheses / buffers   -> one per rank                      (cfg 4)
  * frames of one long integration     -> contiguous blocks of frames per rank, each rank
    accumulates its block from zero, and ONE all-reduce (NCCL over NVLink on GPUs, gloo in
    the CPU tests) sums the weighted partial accumulators                   (cfg 5)
The EMA of src/GUI.jl:175, out <- a*out + (1-a)*m, is linear:
    out_N = a^N out_0 + sum_k (1-a) a^(N-1-k) m_k
so rank g, having run the plain EMA from zero over frames k0..k1-1, only has to scale its
accumulator by a^(N-k1) before the sum.  The result equals the sequential recurrence up to
Float32 rounding order (bit-exactness is a single-GPU property; tolerance stated in the tests).
"""
import numpy as np


def shard_round_robin(n_units, world, rank):
    """unit u -> rank u mod world (independent buffers / hypotheses)"""
    return list(range(rank, n_units, world))


def shard_contiguous(n_units, world, rank):
    """[start, stop) of the contiguous block of rank `rank`; the first n_units % world ranks get one extra"""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def ema_tail_weight(alpha, frames_after):
    """a^(frames_after): weight of a block's accumulator once `frames_after` later frames have been folded in"""
    return float(np.float64(alpha) ** int(frames_after))


class _DevicePtr:
    """expose a raw device pointer to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def accumulator_tensor(chain):
    """the chain's imageOut accumulator (600*800 floats, scan order) as a torch CUDA tensor view"""
    import torch
    ptr, n = chain.accumulator_ptr()
    return torch.as_tensor(_DevicePtr(ptr, n), device=torch.device("cuda", chain.device))


def allreduce_partial(acc, alpha, frames_after, group=None, sum_mode=False, total_frames=None):
    """fold one rank's partial accumulator into the global image (in place, every rank gets the result).
    acc: torch tensor (CUDA with NCCL, CPU with gloo).  EMA mode: scale by a^frames_after, all-reduce(sum).
    Sum mode (TSDR_CHAIN_SUM): all-reduce(sum), then divide by total_frames when given (plain mean)."""
    import torch.distributed as dist
    if not sum_mode:
        acc.mul_(ema_tail_weight(alpha, frames_after))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    if sum_mode and total_frames:
        acc.div_(float(total_frames))
    return acc


def integrate_frames_sharded(make_chain, frame_source, n_frames, alpha, rank, world, comm=None):
    """cfg 5: integrate `n_frames` frames over `world` GPUs.
    make_chain(): a fresh tempestsdr_b200.Chain (zero accumulator) on this rank's GPU, sized for one block.
    frame_source(k0, k1): complex64 samples of frames k0..k1-1 (host array).
    comm: this rank's tempestsdr_b200.Comm (the C-ABI NCCL communicator); None for a single rank.
    Returns (combined image as numpy (600, 800), chain); the image is present on every rank."""
    k0, k1 = shard_contiguous(n_frames, world, rank)
    ch = make_chain()
    if k0 > 0:
        ch.prime(frame_source(k0 - 1, k0))  # halo frame: gives this block the sequential run's first s_y
    if k1 > k0:
        ch.push(frame_source(k0, k1))
    weight = ema_tail_weight(alpha, n_frames - k1)
    if comm is not None and world > 1:
        comm.allreduce_chain(ch, weight)
    elif weight != 1.0:
        ch.scale_accumulator(weight)
    return ch.image(), ch
