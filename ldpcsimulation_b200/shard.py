"""(e) Multi-GPU partitioning: frames are independent, so ranks own disjoint ranges of GLOBAL frame
ids.  Because the channel is a pure function of (seed, frame id), the union of the ranks' runs is
bit-identical to one run over the whole range; the only exchange is the final sum of the counters."""


def shard_range(frame_begin, n_frames, rank, world):
    """Contiguous split [g*F/G, (g+1)*F/G) of n_frames frames starting at frame_begin."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return frame_begin + lo, hi - lo


def step_range(step, rank, world, frames_per_rank):
    """Weak-scaling schedule used by bench.py: step s hands every rank its own block of frames."""
    return (step * world + rank) * frames_per_rank, frames_per_rank


COUNTER_KEYS = ("errors", "uncodedErrors", "totalBits", "totalWords", "wordErrors", "totalIterations",
                "smoothingUsed", "undetectedWords")


def pack_counters(counters, hists=()):
    """Flatten the counter dict (+ optional histogram arrays) into one int64 list for a sum all-reduce."""
    flat = [int(counters[k]) for k in COUNTER_KEYS]
    for h in hists:
        flat.extend(int(x) for x in h)
    return flat


def unpack_counters(flat, hist_lens=()):
    out = {k: int(flat[i]) for i, k in enumerate(COUNTER_KEYS)}
    p, hs = len(COUNTER_KEYS), []
    for n in hist_lens:
        hs.append([int(x) for x in flat[p:p + n]])
        p += n
    return out, hs
