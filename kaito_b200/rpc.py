"""Retrieve RPC between the engine process and the HTTP front-end workers (kaito_b200/frontend.py).

A pod is one replica with all its GPUs (pkg/ragengine/manifests/manifests.go:81) and the engine is one process, but HTTP
parsing, validation and response writing are plain CPU work that does not have to share the engine's interpreter.  With
KRAG_HTTP_WORKERS=N the service starts N front-end processes on the public port (SO_REUSEPORT); each keeps ONE stream
connection to this server and forwards every well-formed POST /retrieve over it:

    frame    = u32 big-endian length | msgpack payload
    request  = [[id, index_name, query, top_k, metadata_filter | nil], ...]
    reply    = [[id, http_status, body_bytes], ...]          (body: the response JSON, or {"detail": ...} for an error)

Requests go straight into the coalescer (kaito_b200/batcher.py) with a sink instead of a Future; the replies of one
coalescing window are packed into one frame per connection and handed to the event loop with a single thread-safe call.
Everything that is not a well-formed /retrieve is proxied by the workers to the engine's own HTTP port, so every other route,
every validation error and /metrics are produced by the same FastAPI app as without workers."""
from __future__ import annotations

import asyncio
import json
import os
import struct
import threading
import time

import msgpack

_LEN = struct.Struct(">I")
MAX_FRAME = 64 << 20


def pack_frame(obj) -> bytes:
    body = msgpack.packb(obj, use_bin_type=True)
    return _LEN.pack(len(body)) + body


async def read_frame(reader: asyncio.StreamReader):
    head = await reader.readexactly(4)
    (n,) = _LEN.unpack(head)
    if n > MAX_FRAME:
        raise ValueError("frame too large")
    return msgpack.unpackb(await reader.readexactly(n), raw=False)


class _Conn:
    __slots__ = ("writer", "pending")

    def __init__(self, writer):
        self.writer, self.pending = writer, []


class RetrieveRpcServer:
    """engine side: a unix-socket (or loopback TCP) server on its own event-loop thread"""

    def __init__(self, batcher, observe, http_exception_types: tuple, path: str | None = None, port: int | None = None):
        self.batcher, self.observe, self.exc = batcher, observe, http_exception_types
        self.path, self.port = path, port
        self._dirty: set[_Conn] = set()
        self._conns: set[_Conn] = set()          # event-loop thread only
        self._mu = threading.Lock()              # sinks and flushes run on the coalescer's dispatcher threads (more than one)
        self._loop = asyncio.new_event_loop()
        self._started = threading.Event()
        self._server = None
        self.requests = 0
        batcher.after_batch = self._flush
        self._t = threading.Thread(target=self._run, name="krag-retrieve-rpc", daemon=True)
        self._t.start()
        self._started.wait(10)

    # ------------------------------------------------------------------ event-loop thread
    def _run(self):
        asyncio.set_event_loop(self._loop)

        async def start():
            if self.path:
                if os.path.exists(self.path):
                    os.unlink(self.path)
                self._server = await asyncio.start_unix_server(self._handle, path=self.path)
                os.chmod(self.path, 0o600)           # the workers are this process's children: nobody else talks to the engine
            else:
                self._server = await asyncio.start_server(self._handle, host="127.0.0.1", port=self.port or 0)
                self.port = self._server.sockets[0].getsockname()[1]
            self._started.set()
        self._loop.run_until_complete(start())
        self._loop.run_forever()

    async def _handle(self, reader, writer):
        conn = _Conn(writer)
        self._conns.add(conn)
        submit = self.batcher.submit_bytes
        try:
            while True:
                for rid, index_name, query, top_k, flt in await read_frame(reader):
                    self.requests += 1
                    submit(index_name, query, top_k, flt, sink=self._sink(conn, rid))
        except (asyncio.IncompleteReadError, ConnectionError, ValueError):
            pass
        finally:
            conn.writer = None
            self._conns.discard(conn)
            try:
                writer.close()
            except Exception:
                pass

    # ------------------------------------------------------------------ coalescer's dispatcher thread
    def _sink(self, conn: _Conn, rid: int):
        t0 = time.perf_counter()

        def deliver(out):
            if isinstance(out, tuple):
                reply = (rid, 200, out[0])
                self.observe("success", time.perf_counter() - t0, {"count": out[1], "results": [{"score": s} for s in out[2]]})
            else:
                status = getattr(out, "status_code", 500) if isinstance(out, self.exc) else 500
                detail = getattr(out, "detail", None) if isinstance(out, self.exc) else str(out)
                reply = (rid, status, json.dumps({"detail": detail}).encode("utf-8"))
                self.observe("failure", time.perf_counter() - t0, None)
            with self._mu:
                conn.pending.append(reply)
                self._dirty.add(conn)
        return deliver

    def _flush(self):
        """after every coalescing window: one frame per connection that got replies"""
        with self._mu:
            dirty, self._dirty = self._dirty, set()
            out = []
            for conn in dirty:
                replies, conn.pending = conn.pending, []
                if conn.writer is not None and replies:
                    out.append((conn.writer, replies))
        for w, replies in out:
            self._loop.call_soon_threadsafe(self._write, w, pack_frame(replies))

    @staticmethod
    def _write(writer, frame):
        try:
            writer.write(frame)
        except Exception:
            pass

    def close(self):
        def stop():
            if self._server is not None:
                self._server.close()
            for conn in list(self._conns):       # the workers see EOF, fail what is in flight and reconnect later
                w, conn.writer = conn.writer, None
                if w is not None:
                    w.close()
            self._loop.call_later(0.05, self._loop.stop)
        self._loop.call_soon_threadsafe(stop)
        self._t.join(timeout=5)
        if self.batcher.after_batch == self._flush:
            self.batcher.after_batch = None
        if self.path and os.path.exists(self.path):
            try:
                os.unlink(self.path)
            except OSError:
                pass


class RetrieveRpcClient:
    """front-end side (asyncio): one connection, requests multiplexed by id"""

    def __init__(self, path: str | None = None, port: int | None = None):
        self.path, self.port = path, port
        self._writer = None
        self._futs: dict[int, asyncio.Future] = {}
        self._next = 0
        self._lock = asyncio.Lock()
        self._reader_task = None

    async def _connect(self):
        async with self._lock:
            if self._writer is not None:
                return
            if self.path:
                reader, writer = await asyncio.open_unix_connection(self.path)
            else:
                reader, writer = await asyncio.open_connection("127.0.0.1", self.port)
            self._writer = writer
            self._reader_task = asyncio.ensure_future(self._read(reader))

    async def _read(self, reader):
        try:
            while True:
                for rid, status, body in await read_frame(reader):
                    f = self._futs.pop(rid, None)
                    if f is not None and not f.done():
                        f.set_result((status, body))
        except Exception as e:                       # engine gone: fail what is in flight, reconnect on the next request
            self._writer = None
            futs, self._futs = self._futs, {}
            for f in futs.values():
                if not f.done():
                    f.set_exception(ConnectionError(f"engine connection lost: {e}"))

    async def retrieve(self, index_name: str, query: str, top_k: int, metadata_filter):
        """-> (http status, body bytes)"""
        if self._writer is None:
            await self._connect()
        self._next += 1
        rid = self._next
        f = asyncio.get_running_loop().create_future()
        self._futs[rid] = f
        self._writer.write(pack_frame([[rid, index_name, query, top_k, metadata_filter]]))
        return await f
