"""CPU check of the ALGORITHM behind the HBM-resident index write side (csrc/index_store.cu).

The device applies a whole batch of indexer.Add calls at once.  It relies on two claims about the reference's
per-server golang-lru + inverted map (approximateprefix/indexer.go:52-83, 105-115, 167-182):

  1. after any sequence of Adds, a server's LRU holds exactly the `cap` most recently added DISTINCT hashes
     (Get never touches recency), so a batch can be applied as "stamp every pair with the sequence number of its
     latest Add, then keep the cap newest per server";
  2. the only pairs that outlive their LRU entry are those a single Add evicts itself -- hash i of a call is evicted by
     that call iff at least `cap` distinct hashes follow its last occurrence -- because the second loop of Add
     (indexer.go:76-83) re-inserts every hash of the call into the inverted map.

`BatchModel` below restates exactly what the kernels do (sequence numbers, newest-cap selection, leak rule, RemovePod)
in plain Python; the oracle's indexer runs the same calls one by one.  No GPU involved."""
from __future__ import annotations

import random

import pytest


class BatchModel:
    def __init__(self, default_cap: int):
        self.default_cap = default_cap
        self.cap = {}          # server -> LRU size (absent = no LRU)
        self.seq = {}          # (hash, server) -> sequence number of the latest Add while in the LRU
        self.in_map = set()    # (hash, server) pairs of hashToPods
        self.next_seq = {}     # server -> next sequence number

    def apply(self, calls):
        """calls: list of (server, hashes, num_gpu_blocks) in the order the reference would run them."""
        leaked = set()
        touched = set()
        for server, hashes, nb in calls:                       # k_store_plan / k_store_offsets / k_store_upsert
            if server not in self.cap:
                c = nb if nb > 0 else self.default_cap
                self.cap[server] = c if c > 0 else 1
            cap = self.cap[server]
            base = self.next_seq.get(server, 1)
            n = len(hashes)
            for i, h in enumerate(hashes):
                self.seq[(h, server)] = base + i               # atomicMax: the latest Add wins
                self.in_map.add((h, server))
                leaked.discard((h, server))
            self.next_seq[server] = base + n
            touched.add(server)
            if n > cap:                                        # k_store_leak
                last = {}
                for i, h in enumerate(hashes):
                    last[h] = i
                is_last = [last[h] == i for i, h in enumerate(hashes)]
                after = 0
                for i in range(n - 1, -1, -1):
                    if is_last[i] and after >= cap:
                        leaked.add((hashes[i], server))
                    after += 1 if is_last[i] else 0
        for server in touched:                                 # k_evict_*: keep the cap newest live pairs
            live = sorted(((s, h) for (h, sv), s in self.seq.items() if sv == server), reverse=True)
            for s, h in live[self.cap[server]:]:
                del self.seq[(h, server)]
                if (h, server) not in leaked:                  # the eviction callback prunes the map ...
                    self.in_map.discard((h, server))           # ... unless the same call re-inserted it (leak)

    def remove_pod(self, server):
        if server not in self.cap:
            return
        for (h, sv) in [k for k in self.seq if k[1] == server]:
            del self.seq[(h, sv)]
            self.in_map.discard((h, sv))
        del self.cap[server]

    def get(self, h):
        return {sv for (hh, sv) in self.in_map if hh == h}


@pytest.mark.parametrize("seed,default_cap,universe,n_srv", [(1, 5, 40, 4), (2, 1, 12, 3), (3, 16, 200, 9), (4, 3, 30, 2)])
def test_batched_adds_equal_sequential_adds(orc, seed, default_cap, universe, n_srv):
    rng = random.Random(seed)
    ix = orc.Indexer(default_cap)
    model = BatchModel(default_cap)
    for rnd in range(40):
        calls = []
        for _ in range(rng.choice([1, 3, 25])):
            server = rng.randrange(n_srv)
            n = rng.choice([0, 1, 2, 4, 9, 20])
            base = rng.randrange(universe)
            hashes = [(base + (j if rng.random() < 0.7 else rng.randrange(universe))) % universe + 1 for j in range(n)]
            nb = rng.choice([0, 0, 2, 7])
            calls.append((server, hashes, nb))
            ix.add(hashes, server, nb)
        model.apply(calls)
        if rng.random() < 0.15:
            server = rng.randrange(n_srv)
            ix.remove_pod(server)
            model.remove_pod(server)
        for h in range(1, universe + 1):
            assert model.get(h) == ix.get(h), (seed, rnd, h)
        assert len(model.in_map) == len(ix.export()[0])
