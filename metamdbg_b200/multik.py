"""The multi-k loop of the hot path, on the device-resident minimizer store.

metaMDBG re-runs its `graph` stage for k = firstK, firstK+1, ... on the SAME minimizer-space reads
(src/pipeline/AssemblyPipeline.hpp drives it; src/graph/CreateMdbg.cpp:386-468 is the k > firstK branch):

  * k = firstK      KminmerCounter::execute   (CreateMdbg.hpp:3591-3883)  -> occurrence counts, abundance >= 2
                    [+ rescueKminmers, CreateMdbg.hpp:4517-4640, in default mode]
  * k > firstK      abundance of a k-min-mer = min over its two (k-1)-min-mers of the previous k's table
                    (getRefinedAbundance / IndexKminmerFunctor, CreateMdbg.hpp:3933-4005, 951-1465)

Upstream patches the previous-k table with refined unitig abundances from the contig stage between two k's; that
stage is outside this engine (mdbg_prev_load takes such patches), so the loop below is the part that can stay on the
device: every call is a call into libmdbg_b200.so, nothing is computed in Python.  With more than one rank the
three collective calls (merge, rescue, previous-k replication) are made by every rank in the same order.
"""
from __future__ import annotations

import time


def multi_k_sweep(engine, first_k: int = 4, last_k: int = 21, min_abundance: int = 2, rescue: bool = False,
                  merge: bool = False, on_table=None, synchronize: bool = True, world: int = 1,
                  table_headroom: float = 1.3) -> list[dict]:
    """Count k = first_k on the engine's store, then derive k = first_k+1 .. last_k from the previous table.

    merge: False (one context), True (owner merge with the k-min-mer vectors after every k) or "hashes" (k > first_k:
    mdbg_count_merge_hashes -- (hash, abundance) records only, the light collective of a pure table loop).

    on_table(k, engine) is called while the table of k is current (e.g. to finalize it into kminmerData files);
    returns one dict per k: k, seconds (host wall clock around the device work of that k) and the table statistics.

    Table sizing: the first table is sized for the worst case (every window distinct).  A k > first_k table holds
    k-min-mers whose two (k-1)-min-mers are both in the previous table, i.e. about as many entries as that table, so
    it is sized for table_headroom x (previous entries x world) instead -- the per-k clear / statistics / previous-k
    passes then touch megabytes, not the gigabytes of the worst-case table.  The next-k pass is idempotent, so if
    such a table does fill up (MDBG_ERR_TABLE_FULL) the k is simply redone with the worst-case size.
    """
    from .engine import MdbgError
    out = []
    n_prev_entries = 0
    for k in range(first_k, last_k + 1):
        if synchronize:
            engine.synchronize()
        t0 = time.perf_counter()
        if k == first_k:
            engine.count_begin(k, 0)
            engine.count_add_store()
            if merge:
                engine.count_merge()
            n_rescued_reads = engine.count_rescue() if rescue else 0
        else:
            engine.prev_from_current(min_abundance)
            expected = int(table_headroom * max(n_prev_entries, 256) * max(1, world)) if table_headroom > 0 else 0
            try:
                engine.count_begin(k, expected)
                engine.count_add_store_next_k()
            except MdbgError as e:
                if e.status != 4 or expected == 0:            # 4 = MDBG_ERR_TABLE_FULL
                    raise
                engine.count_begin(k, 0)                      # worst-case size; the pass is idempotent
                engine.count_add_store_next_k()
            if merge == "hashes":                             # keys + abundances only: no vectors on the owners
                engine.count_merge_hashes()
            elif merge:
                engine.count_merge()
            n_rescued_reads = 0
        stats = engine.count_stats(min_abundance)         # device-side reduction + one small D2H: ends the k's work
        dt = time.perf_counter() - t0
        if on_table is not None:
            on_table(k, engine)
        n_prev_entries = stats["n_entries"]
        out.append(dict(k=k, seconds=dt, n_entries=stats["n_entries"], checksum=stats["checksum"],
                        n_reads_rescued=n_rescued_reads))
    return out
