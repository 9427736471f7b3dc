"""Fast path of POST /retrieve below FastAPI's routing stack (kaito_b200.service installs it as an ASGI middleware).

The reference's handler (main.py:742-771) is a pydantic-validated FastAPI route; at several thousand requests per second the
routing / dependency / validation / response-model machinery of FastAPI costs more event-loop time than the whole retrieval.
This middleware answers well-formed requests itself -- body -> json -> coalescer future -> bytes -- and hands EVERYTHING else
(any other route, malformed JSON, missing or out-of-range fields) to the FastAPI app untouched, so validation errors keep the
exact 422 bodies FastAPI produces."""
from __future__ import annotations

import asyncio
import json
import time


class FastRetrieve:
    def __init__(self, inner, submit, observe, max_top_k: int, http_exception_types: tuple):
        """submit(index_name, query, top_k, metadata_filter) -> concurrent.futures.Future of the response dict;
        observe(status: 'success'|'failure', seconds, out_or_None): the route's Prometheus bookkeeping"""
        self.inner, self.submit, self.observe, self.max_top_k, self.exc = inner, submit, observe, max_top_k, http_exception_types

    async def __call__(self, scope, receive, send):
        if scope["type"] != "http" or scope["method"] != "POST" or scope["path"] != "/retrieve":
            return await self.inner(scope, receive, send)
        chunks, more = [], True
        while more:
            msg = await receive()
            if msg["type"] != "http.request":        # disconnect
                return
            chunks.append(msg.get("body", b""))
            more = msg.get("more_body", False)
        body = b"".join(chunks)
        req = self._parse(body)
        if req is None:                               # let FastAPI produce its own 4xx for this body
            sent = False

            async def replay():
                nonlocal sent
                if sent:
                    return await receive()
                sent = True
                return {"type": "http.request", "body": body, "more_body": False}
            return await self.inner(scope, replay, send)
        t0 = time.perf_counter()
        status, payload = 200, None
        try:
            out = await asyncio.wrap_future(self.submit(*req))
            if isinstance(out, tuple):                # (json bytes, count, scores) from VectorStore.retrieve_batch_bytes
                payload = out[0]
                self.observe("success", time.perf_counter() - t0, {"count": out[1], "results": [{"score": s} for s in out[2]]})
            else:
                for r in out["results"]:              # models.NodeWithScore: the optional scores serialise as null when unset
                    r.setdefault("dense_score", None); r.setdefault("sparse_score", None); r.setdefault("source", None)
                    r.setdefault("metadata", None)
                payload = json.dumps(out, ensure_ascii=False, allow_nan=False, separators=(",", ":")).encode("utf-8")
                self.observe("success", time.perf_counter() - t0, out)
        except self.exc as e:                         # vs.HTTPException / fastapi.HTTPException: {"detail": ...}
            status, payload = e.status_code, json.dumps({"detail": e.detail}).encode("utf-8")
            self.observe("failure", time.perf_counter() - t0, None)
        except Exception as e:
            status, payload = 500, json.dumps({"detail": str(e)}).encode("utf-8")
            self.observe("failure", time.perf_counter() - t0, None)
        await send({"type": "http.response.start", "status": status,
                    "headers": [(b"content-type", b"application/json"), (b"content-length", str(len(payload)).encode())]})
        await send({"type": "http.response.body", "body": payload})

    def _parse(self, body: bytes):
        """(index_name, query, max_node_count, metadata_filter) for a request RetrieveRequest would accept, else None"""
        try:
            d = json.loads(body)
        except Exception:
            return None
        if not isinstance(d, dict):
            return None
        name, query, k, flt = d.get("index_name"), d.get("query"), d.get("max_node_count", 5), d.get("metadata_filter")
        if not isinstance(name, str) or not isinstance(query, str):
            return None
        if isinstance(k, bool) or not isinstance(k, int) or k < 1 or k > self.max_top_k:
            return None
        if flt is not None and not isinstance(flt, dict):
            return None
        return name, query, k, flt
