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
    def __init__(self, store, max_batch: int | None = None, max_wait_s: float | None = None, dispatchers: int | None = None):
        self.store = store
        self.max_batch = int(os.getenv("KRAG_MAX_BATCH", "256")) if max_batch is None else max_batch
        self.max_wait_s = float(os.getenv("KRAG_BATCH_WINDOW_US", "200")) * 1e-6 if max_wait_s is None else max_wait_s
        self._q: queue.Queue = queue.Queue()
        self._stop = False
        self.batches = 0                  # engine calls issued
        self.requests = 0                 # requests answered
        self.max_seen = 0                 # largest group sent to the engine
        self.after_batch = None           # optional hook, called on the dispatcher thread after every window (RPC flush)
        # two dispatcher threads: while one batch is inside the engine call (GIL released) the other collects, tokenises and
        # later serialises the next one, so the per-batch host work overlaps the GPU instead of adding to it
        self._stats_mu = threading.Lock()
        n = max(1, int(os.getenv("KRAG_BATCH_DISPATCHERS", "2")) if dispatchers is None else dispatchers)
        self._ts = [threading.Thread(target=self._run, name=f"krag-retrieve-batcher-{i}", daemon=True) for i in range(n)]
        for t in self._ts:
            t.start()

    @property
    def enabled(self) -> bool:
        return self.max_wait_s > 0 and self.max_batch > 1

    def submit(self, index_name: str, query: str, top_k: int, metadata_filter: dict | None) -> Future:
        f: Future = Future()
        self._q.put((index_name, query, top_k, metadata_filter, f, False))
        return f

    def submit_bytes(self, index_name: str, query: str, top_k: int, metadata_filter: dict | None, sink=None):
        """like submit(), answered with (json bytes, count, fused scores) from VectorStore.retrieve_batch_bytes.
        sink=None: returns a Future.  sink=callable: no Future is created; sink(result_or_exception) is called on the
        dispatcher thread (the RPC server of the front-end workers passes a closure that queues the reply frame)."""
        f = Future() if sink is None else sink
        self._q.put((index_name, query, top_k, metadata_filter, f, True))
        return f if sink is None else None

    def retrieve(self, index_name: str, query: str, top_k: int, metadata_filter: dict | None):
        """blocking convenience wrapper (sync route handlers)"""
        return self.submit(index_name, query, top_k, metadata_filter).result()

    def close(self):
        self._stop = True
        for _ in self._ts:
            self._q.put(None)
        for t in self._ts:
            t.join(timeout=5)

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
                    self._q.put(None)                       # the sentinel was meant for one thread: pass it on
                    break
                batch.append(item)
            groups: dict[tuple, list] = {}
            for it in batch:
                key = (it[0], it[2], json.dumps(it[3], sort_keys=True, default=str) if it[3] else None, it[5])
                groups.setdefault(key, []).append(it)
            for (index_name, top_k, _, as_bytes), items in groups.items():
                with self._stats_mu:
                    self.batches += 1
                    self.requests += len(items)
                    self.max_seen = max(self.max_seen, len(items))
                fn = self.store.retrieve_batch_bytes if as_bytes else self.store.retrieve_batch
                try:
                    outs = fn(index_name, [it[1] for it in items], top_k, items[0][3])
                except Exception as e:                       # the whole group failed the same way (404, engine error)
                    outs = [e] * len(items)
                for it, out in zip(items, outs):
                    self._deliver(it[4], out)
            if self.after_batch is not None:
                self.after_batch()

    @staticmethod
    def _deliver(target, out):
        if isinstance(target, Future):
            if isinstance(out, Exception):
                target.set_exception(out)
            else:
                target.set_result(out)
        else:
            try:
                target(out)
            except Exception:                                # a sink must not take the dispatcher down
                pass
