"""HTTP front-end worker of the RAGEngine service (KRAG_HTTP_WORKERS=N; see kaito_b200/rpc.py for the why).

    python -m kaito_b200.frontend --port 5000 --rpc-path /tmp/krag-rpc.sock --engine-http 127.0.0.1:5001

N of these share the public port through SO_REUSEPORT.  A worker answers well-formed POST /retrieve requests itself
(body -> json -> RPC to the engine's coalescer -> the engine's response bytes, untouched) and reverse-proxies everything else
-- every other route, malformed /retrieve bodies, /metrics, /health -- to the engine process's own HTTP port, so behaviour,
error bodies and metrics are those of the single-process service (presets/ragengine/main.py)."""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import socket
import sys

from .rpc import RetrieveRpcClient

HOP_BY_HOP = {b"connection", b"keep-alive", b"proxy-authenticate", b"proxy-authorization", b"te", b"trailers", b"transfer-encoding",
              b"upgrade", b"host", b"content-length"}


def parse_retrieve(body: bytes, max_top_k: int):
    """(index_name, query, max_node_count, metadata_filter) for a body models.RetrieveRequest accepts, else None (the engine's
    FastAPI app then produces its own 4xx).  Same rule as kaito_b200.fast_retrieve.FastRetrieve._parse."""
    try:
        d = json.loads(body)
    except Exception:
        return None
    if not isinstance(d, dict):
        return None
    name, query, k, flt = d.get("index_name"), d.get("query"), d.get("max_node_count", 5), d.get("metadata_filter")
    if not isinstance(name, str) or not isinstance(query, str):
        return None
    if isinstance(k, bool) or not isinstance(k, int) or k < 1 or k > max_top_k:
        return None
    if flt is not None and not isinstance(flt, dict):
        return None
    return name, query, k, flt


class FrontApp:
    def __init__(self, rpc: RetrieveRpcClient, engine_http: str, max_top_k: int):
        self.rpc, self.engine_http, self.max_top_k = rpc, engine_http, max_top_k
        self.timeout_s = float(os.getenv("KRAG_RPC_TIMEOUT_S", "300"))
        self._session = None

    async def __call__(self, scope, receive, send):
        if scope["type"] == "lifespan":
            while True:
                msg = await receive()
                if msg["type"] == "lifespan.startup":
                    await send({"type": "lifespan.startup.complete"})
                elif msg["type"] == "lifespan.shutdown":
                    if self._session is not None:
                        await self._session.close()
                    await send({"type": "lifespan.shutdown.complete"})
                    return
        if scope["type"] != "http":
            return
        chunks, more = [], True
        while more:
            msg = await receive()
            if msg["type"] != "http.request":
                return
            chunks.append(msg.get("body", b""))
            more = msg.get("more_body", False)
        body = chunks[0] if len(chunks) == 1 else b"".join(chunks)
        if scope["method"] == "POST" and scope["path"] == "/retrieve":
            req = parse_retrieve(body, self.max_top_k)
            if req is not None:
                try:
                    status, payload = await asyncio.wait_for(self.rpc.retrieve(*req), timeout=self.timeout_s)
                except (ConnectionError, OSError) as e:
                    status, payload = 503, json.dumps({"detail": f"engine unreachable: {e}"}).encode("utf-8")
                except asyncio.TimeoutError:
                    status, payload = 504, json.dumps({"detail": "engine did not answer in time"}).encode("utf-8")
                await send({"type": "http.response.start", "status": status,
                            "headers": [(b"content-type", b"application/json"), (b"content-length", str(len(payload)).encode())]})
                await send({"type": "http.response.body", "body": payload})
                return
        await self._proxy(scope, body, send)

    async def _proxy(self, scope, body: bytes, send):
        import aiohttp
        if self._session is None:
            self._session = aiohttp.ClientSession(auto_decompress=False, timeout=aiohttp.ClientTimeout(total=None))
        qs = scope.get("query_string", b"")
        url = f"http://{self.engine_http}{scope['path']}" + (("?" + qs.decode("latin-1")) if qs else "")
        headers = [(k.decode("latin-1"), v.decode("latin-1")) for k, v in scope["headers"] if k.lower() not in HOP_BY_HOP]
        try:
            async with self._session.request(scope["method"], url, data=body if body else None, headers=headers, allow_redirects=False) as r:
                payload = await r.read()
                out_headers = [(k.lower().encode("latin-1"), v.encode("latin-1")) for k, v in r.headers.items()
                               if k.lower().encode("latin-1") not in HOP_BY_HOP]
                out_headers.append((b"content-length", str(len(payload)).encode()))
                await send({"type": "http.response.start", "status": r.status, "headers": out_headers})
                await send({"type": "http.response.body", "body": payload})
        except aiohttp.ClientError as e:
            payload = json.dumps({"detail": f"engine unreachable: {e}"}).encode("utf-8")
            await send({"type": "http.response.start", "status": 503,
                        "headers": [(b"content-type", b"application/json"), (b"content-length", str(len(payload)).encode())]})
            await send({"type": "http.response.body", "body": payload})


def listen_socket(host: str, port: int) -> socket.socket:
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEPORT, 1)      # the kernel spreads new connections over the workers
    s.bind((host, port))
    s.listen(2048)
    s.setblocking(False)
    return s


def spawn(n: int, host: str, port: int, engine_http: str, rpc_path: str | None, rpc_port: int | None, max_top_k: int) -> list:
    """start n front-end workers (subprocess.Popen objects); the caller terminates them"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "kaito_b200.frontend", "--host", host, "--port", str(port), "--engine-http", engine_http,
           "--max-top-k", str(max_top_k)]
    cmd += ["--rpc-path", rpc_path] if rpc_path else ["--rpc-port", str(rpc_port)]
    return [subprocess.Popen(cmd, env=env) for _ in range(n)]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=5000)
    ap.add_argument("--engine-http", default="127.0.0.1:5001")
    ap.add_argument("--rpc-path", default=None)
    ap.add_argument("--rpc-port", type=int, default=None)
    ap.add_argument("--max-top-k", type=int, default=300)
    a = ap.parse_args(argv)
    import uvicorn
    app = FrontApp(RetrieveRpcClient(a.rpc_path, a.rpc_port), a.engine_http, a.max_top_k)
    server = uvicorn.Server(uvicorn.Config(app, log_level="warning", access_log=False, lifespan="on"))
    server.run(sockets=[listen_socket(a.host, a.port)])


if __name__ == "__main__":
    main()
