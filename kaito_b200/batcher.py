"""Request coalescer in front of POST /retrieve.

The reference serves one query per request (presets/ragengine/main.py:742-771) and a pod is exactly one replica
(pkg/ragengine/manifests/manifests.go:81): everything the GPU(s) of the pod can do has to come out of that one process.
A batch of 256 queries costs the engine barely more than a single one (the corpus is streamed from HBM once per batch), so
concurrent requests are collected for a short window and answered by ONE engine call:

    request threads / tasks --submit()--> queue --dispatcher thread--> VectorStore.retrieve_batch(group) --> futures

The window closes when `max_batch` requests are waiting or `max_wait_s` after the first one arrived (default 200 us:
well below the engine's own batch time, so a lone request is not noticeably delayed).  Requests are grouped by
(index, top_k, metadata filter); each group is one engine call.  KRAG_BATCH_WINDOW_US=0 disables coalescing.
"""
from __future__ import annotations

import json
import os
import queue
import threading
import time
from concurrent.futures import Future


class RetrieveBatcher:
    def __init__(self, store, max_batch: int | None = None, max_wait_s: float | None = None):
        self.store = store
        self.max_batch = int(os.getenv("KRAG_MAX_BATCH", "256")) if max_batch is None else max_batch
        self.max_wait_s = float(os.getenv("KRAG_BATCH_WINDOW_US", "200")) * 1e-6 if max_wait_s is None else max_wait_s
        self._q: queue.Queue = queue.Queue()
        self._stop = False
        self.batches = 0                  # engine calls issued
        self.requests = 0                 # requests answered
        self.max_seen = 0                 # largest group sent to the engine
        self._t = threading.Thread(target=self._run, name="krag-retrieve-batcher", daemon=True)
        self._t.start()

    @property
    def enabled(self) -> bool:
        return self.max_wait_s > 0 and self.max_batch > 1

    def submit(self, index_name: str, query: str, top_k: int, metadata_filter: dict | None) -> Future:
        f: Future = Future()
        self._q.put((index_name, query, top_k, metadata_filter, f))
        return f

    def retrieve(self, index_name: str, query: str, top_k: int, metadata_filter: dict | None):
        """blocking convenience wrapper (sync route handlers)"""
        return self.submit(index_name, query, top_k, metadata_filter).result()

    def close(self):
        self._stop = True
        self._q.put(None)
        self._t.join(timeout=5)

    # ------------------------------------------------------------------ dispatcher
    def _run(self):
        while not self._stop:
            first = self._q.get()
            if first is None:
                break
            batch = [first]
            deadline = time.perf_counter() + self.max_wait_s
            while len(batch) < self.max_batch:
                try:                                        # drain what is already there without sleeping
                    item = self._q.get_nowait()
                except queue.Empty:
                    remaining = deadline - time.perf_counter()
                    if remaining <= 0:
                        break
                    try:
                        item = self._q.get(timeout=remaining)
                    except queue.Empty:
                        break
                if item is None:
                    self._stop = True
                    break
                batch.append(item)
            groups: dict[tuple, list] = {}
            for it in batch:
                key = (it[0], it[2], json.dumps(it[3], sort_keys=True, default=str) if it[3] else None)
                groups.setdefault(key, []).append(it)
            for (index_name, top_k, _), items in groups.items():
                self.batches += 1
                self.requests += len(items)
                self.max_seen = max(self.max_seen, len(items))
                try:
                    outs = self.store.retrieve_batch(index_name, [it[1] for it in items], top_k, items[0][3])
                    for it, out in zip(items, outs):
                        if isinstance(out, Exception):
                            it[4].set_exception(out)
                        else:
                            it[4].set_result(out)
                except Exception as e:                       # the whole group failed the same way (404, engine error)
                    for it in items:
                        if not it[4].done():
                            it[4].set_exception(e)
