"""Frame sharding across GPUs: frames are independent (intra-only analysis), so frame `poc` simply goes
to GPU `poc mod G`; there is no data-path collective -- the host reassembles results in poc order.
Mirrors the CLI's --NumGpus worker loop (csrc/main.cpp:gpu_worker)."""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """POCs owned by `rank` of `world` (round robin, like main.cpp: next = g, g + G, ...)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def merge_in_poc_order(per_rank: Sequence[Dict[int, object]], n_frames: int) -> List[object]:
    """Reassemble {poc: result} dicts of all ranks into a poc-ordered list; every poc exactly once."""
    out: List[object] = [None] * n_frames
    seen = set()
    for d in per_rank:
        for poc, r in d.items():
            if poc in seen or not 0 <= poc < n_frames:
                raise ValueError(f"poc {poc} delivered twice or out of range")
            seen.add(poc)
            out[poc] = r
    if len(seen) != n_frames:
        raise ValueError(f"missing pocs: {sorted(set(range(n_frames)) - seen)}")
    return out


def run_pipelined(engine, frames: Sequence, pocs: Sequence[int], consume: Callable[[int, object], None]) -> None:
    """Drive one engine over its shard with all slots in flight (submit until full, collect FIFO)."""
    slots = engine.cfg.slots
    sub = got = 0
    while got < len(pocs):
        while sub < len(pocs) and engine.in_flight() < slots:
            buf = engine.next_input()
            buf[...] = frames[pocs[sub]]
            engine.submit(buf, poc=pocs[sub])
            sub += 1
        r = engine.collect()
        consume(r.poc, r)
        got += 1
