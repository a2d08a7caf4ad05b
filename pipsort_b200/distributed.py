"""One process per GPU: shard one locus over the ranks of a torch.distributed group.

The configurations of a locus are independent; the only coupling is the final log-sum-exp.  So

* exhaustive (postcal.cpp:716-1092): the union-subset rank space [0, sum_j C(U,j)) is cut into `world`
  contiguous, work-weighted ranges (pipsort_shard_ranks); LD / z / maps are replicated; every rank runs the same
  single launch on its range; the accumulator stores -- plain doubles whose element-wise SUM is the accumulator
  state of the union, configuration count and error counters included -- are combined either by the engine's own
  peer-memory kernels over NVLink (connect_p2p + collective="p2p") or by ONE all-reduce(sum).
* stochastic shotgun search (sss_postcal.cpp:223-255): every iteration's list of unseen neighbours is cut into
  `world` contiguous slices, the per-neighbour scores are all-gathered (the sampling step needs all of them on every
  rank, sss_postcal.cpp:289-343), the accumulators stay rank-partial until they are read.

Nothing here computes: the engine does (CUDA), torch.distributed moves the small vectors (NCCL on GPUs; the host
logic is exercised with gloo in tests/test_sharding_cpu.py through an engine stand-in).
"""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world_and_rank(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


CUDA_STREAM_LEGACY = 1     # cudaStreamLegacy: the handle that NAMES the default stream (0 means "the engine's own stream" to set_stream)


def _torch_stream_handle():
    """torch's current CUDA stream as a handle pipsort_set_stream understands.  torch's default stream is handle 0, which the
    C-ABI reads as "back to the engine's own (non-blocking) stream" -- an engine "bound" that way is not ordered with NCCL at
    all; the default stream is therefore passed by its explicit name, cudaStreamLegacy."""
    import torch
    return torch.cuda.current_stream().cuda_stream or CUDA_STREAM_LEGACY


def bind_engine_to_current_stream(engine):
    """Make the engine launch on torch's current CUDA stream: the NCCL all-reduce torch enqueues is then ordered
    after the engine's kernels (and the finalize kernel after the all-reduce) without host synchronisation."""
    engine.set_stream(_torch_stream_handle())


def _order_with_torch(engine):
    """Stream ordering between the engine's launches and what torch enqueues (NCCL collectives, clone / copy_): the
    engine normally works on its OWN non-blocking stream, torch on its current stream, and nothing orders one after the
    other -- an all-reduce could read the accumulator store before the exhaustive kernel has finished, and finalize
    could run before the all-reduce has.  Every helper below that mixes the two therefore moves the engine onto torch's
    current stream first (pipsort_set_stream waits for the work already queued on the old stream).  A no-op when the
    caller has done so already (bind_engine_to_current_stream) or when there is no CUDA device (gloo tests)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return
        cur = _torch_stream_handle()
    except Exception:
        return
    if hasattr(engine, "stream") and hasattr(engine, "set_stream") and engine.stream() != cur:
        engine.set_stream(cur)


def slice_bounds(n, world):
    """Contiguous near-equal slices of n items: bounds[r] .. bounds[r+1] belongs to rank r."""
    return [(n * r) // world for r in range(world + 1)]


_handle_cache = {}


def connect_p2p(engine, group=None, root=0):
    """Exchange the engines' mailbox handles (CUDA IPC) over the group and map every peer: after this the combine step
    can run over NVLink peer memory (collective="p2p") instead of NCCL.  Returns False when IPC mapping is not possible
    (the caller then stays with the all-reduce).

    Mailboxes belong to the process, not to the engine (the library hands the same one to the next engine that fits), so
    the table of handles is exchanged once and reused while this rank's handle stays the same -- a run that creates one
    engine per locus pays the all_gather_object once, not per locus.  (All ranks run the same sequence of loci, so they
    all hit or all miss.)"""
    world, rank = world_and_rank(group)
    if world == 1:
        return False
    try:
        mine = engine.p2p_export(world)
    except Exception:
        mine = b""
    key = (id(group), world, rank, root)
    cached = _handle_cache.get(key)
    if cached == "unavailable":
        return False                       # every rank recorded the same verdict (it was agreed below)
    if mine and cached is not None and cached[rank] == mine:
        engine.p2p_connect(cached, rank, root)
        return True
    handles = [None] * world
    _dist().all_gather_object(handles, mine, group=group)
    ok = all(len(h) > 0 for h in handles)
    if ok:
        try:
            engine.p2p_connect(handles, rank, root)
        except Exception:
            ok = False
    flags = [None] * world
    _dist().all_gather_object(flags, ok, group=group)
    ok = all(flags)
    _handle_cache[key] = handles if ok else "unavailable"
    return ok


def run_exhaustive_sharded(engine, c, group=None, bounds=None, collective="allreduce"):
    """reset + this rank's share of computeTotalLikelihood + the combine step.  collective="allreduce": ONE all-reduce
    (sum) of the accumulator store, every rank then holds the whole result.  collective="p2p" (after connect_p2p): the
    non-root ranks add their stores into the root's memory over NVLink, only the root holds the result.
    Asynchronous on the engine's stream; follow with engine.read()."""
    world, rank = world_and_rank(group)
    if collective == "p2p" and world > 1:
        if bounds is None:
            bounds = engine.shard_ranks(c, world)
        engine.reset()
        engine.run_exhaustive(c, bounds[rank], bounds[rank + 1])
        engine.p2p_reduce_to_root()
        return bounds
    if bounds is None:
        bounds = engine.shard_ranks(c, world)
    if len(bounds) != world + 1:
        raise ValueError("need world+1 shard bounds")
    if world > 1:
        _order_with_torch(engine)
    engine.reset()
    engine.run_exhaustive(c, bounds[rank], bounds[rank + 1])
    if world > 1:
        _dist().all_reduce(engine.accumulator_tensor(), group=group)
    return bounds


def pass_exhaustive_sharded(engine, c, bounds, collective="p2p", root=0, group=None):
    """One REPEATABLE pass on accumulators that are already empty (a fresh engine, or the previous pass of this function):
    this rank's share of computeTotalLikelihood, the combine step and the finalize, with every reset folded into the
    kernels that read the store last -- the non-root ranks' push empties their stores, the root's finalize (which also
    sums the peers' stores, straight from their mailbox slots) empties its own.  Two launches on every rank; asynchronous; the root's results are then
    read with engine.fetch().  collective="allreduce" keeps an explicit reset (NCCL reads and writes the store itself)."""
    world, rank = world_and_rank(group)
    if world == 1:
        engine.run_exhaustive(c, bounds[0], bounds[1])
        engine.finalize(reset=True)
        return
    if collective == "p2p":
        engine.run_exhaustive(c, bounds[rank], bounds[rank + 1])
        engine.p2p_combine_finalize()         # non-root: send + empty; root: sum over the GPUs inside the finalize
        return
    _order_with_torch(engine)
    engine.run_exhaustive(c, bounds[rank], bounds[rank + 1])
    _dist().all_reduce(engine.accumulator_tensor(), group=group)
    engine.finalize(reset=True)


def compute_total_likelihood_sharded(engine, c=None, group=None, bounds=None, collective="allreduce", root=0):
    """PostCal::computeTotalLikelihood (postcal.cpp:716) with the rank space sharded over the group.  With
    collective="p2p" only the root returns Results (the other ranks return None)."""
    c = engine.max_causal if c is None else c
    run_exhaustive_sharded(engine, c, group, bounds, collective)
    world, rank = world_and_rank(group)
    if collective == "p2p" and world > 1 and rank != root:
        engine.sync()
        return None
    return engine.read()


def score_union_configs_sharded(engine, idx, make_updates=None, group=None):
    """One SSS neighbourhood (sss_postcal.cpp:223-255) split over the group: rank r scores slice r with accumulator
    updates, the max-|l| values of all slices are all-gathered.  Returns max_abs_l[n] (identical on every rank)."""
    world, rank = world_and_rank(group)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    n = idx.shape[0]
    mu = None if make_updates is None else np.ascontiguousarray(make_updates, dtype=np.uint8)
    if world == 1:
        return engine.score_union_configs(idx, mu)
    import torch
    _order_with_torch(engine)
    b = slice_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    mine = engine.score_union_configs(idx[lo:hi], None if mu is None else mu[lo:hi]) if hi > lo else np.zeros(0)
    width = max(b[r + 1] - b[r] for r in range(world))
    dev = engine.accumulator_tensor().device
    send = torch.zeros(width, dtype=torch.float64, device=dev)
    send[:hi - lo] = torch.from_numpy(mine).to(dev)
    recv = torch.empty(world * width, dtype=torch.float64, device=dev)
    _dist().all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, width)
    return np.concatenate([recv[r, :b[r + 1] - b[r]] for r in range(world)])


def sss_sharded(engine, c=None, max_iterations=1000, group=None, root=0):
    """sss_computeTotalLikelihood (sss_postcal.cpp:102-380) with every round's neighbourhood split over the ranks
    (pipsort_sss_sharded; connect_p2p first).  Returns (Results on the root / None elsewhere, iterations, stop_reason)."""
    world, rank = world_and_rank(group)
    engine.reset()
    if world == 1:
        return engine.sss(c, max_iterations)
    it, why = engine.sss_sharded(c, max_iterations)
    engine.p2p_reduce_to_root()
    if rank != root:
        engine.sync()
        return None, it, why
    return engine.read(), it, why


def read_sharded(engine, group=None):
    """Results of rank-partial accumulators (after score_union_configs_sharded): all-reduce a COPY of the store so the
    partial sums can keep accumulating, finalize from the copy, restore."""
    world, _ = world_and_rank(group)
    if world == 1:
        return engine.read()
    _order_with_torch(engine)
    acc = engine.accumulator_tensor()
    keep = acc.clone()
    _dist().all_reduce(acc, group=group)
    res = engine.read()
    acc.copy_(keep)
    return res
